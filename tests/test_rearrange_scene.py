"""Per-reset heterogeneous block scenes as one padded batch (robogym_b200/rearrange_scene.py): the reference
recompiles its model at every reset with another number of blocks, another block size / scale and another material
(robogym/envs/rearrange/common/base.py:850-856,897-906); here one model with the maximum number of blocks carries per-environment
rows.  Checks: the rows `BatchedBlockScene` writes equal the arrays of models COMPILED with those sizes; the kernel logic (CPU
emulation) stepping a model patched with those rows matches the oracle on the separately compiled model; (gpu) one CUDA batch
with different blocks / materials / block counts per environment against one oracle per environment."""
import os
import sys

import numpy as np
import pytest

import pyemu
from helpers import oracle_pair
from robogym_b200 import mjcf, modelblob

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "stubs"))
NOBJ = 4
TABLE_TOP = 0.453 + 0.03324


def scene_xml(half=(0.0254,) * NOBJ, density=1000.0, friction=(1.0, 0.005, 0.0001), solref=(0.02, 1.0)):
    """a table, a floor and NOBJ free blocks (condim 6, elliptic cones, the rearrange material's margin)"""
    body = ""
    for k in range(NOBJ):
        h = half[k] if np.ndim(half[k]) else (half[k],) * 3
        body += (f'<body name="object{k}" pos="{0.1 * k} 0 {TABLE_TOP + h[2]}"><joint name="object{k}:joint" type="free" damping="0.01" armature="0.001"/>'
                 f'<geom name="object{k}" type="box" size="{h[0]} {h[1]} {h[2]}" density="{density}" condim="6" margin="0.00005" '
                 f'friction="{friction[0]} {friction[1]} {friction[2]}" solref="{solref[0]} {solref[1]}"/></body>\n')
    return f"""<mujoco><compiler angle="radian" coordinate="local"/>
<option timestep="0.002" iterations="50" tolerance="1e-10" cone="elliptic" impratio="10"/><size nuserdata="0" njmax="500" nconmax="100"/>
<worldbody>
<body name="floor" pos="0 0 0"><geom name="floor" type="plane" size="5 5 1" condim="3"/></body>
<body name="table" pos="0.15 0 0.453"><geom name="table" type="box" size="0.4 0.3 0.03324" condim="3"/></body>
{body}</worldbody></mujoco>"""


class _RecordingSim:
    """just enough of BatchedSim for BatchedBlockScene on the host: records set_param rows, runs mjcf.set_const per environment"""

    def __init__(self, cm, nenv):
        import torch

        from oracle_generic_sim import _Model

        self.torch = torch
        self.model = _Model(cm.blob())
        self.nenv = nenv
        self.qpos = torch.tensor(np.array(cm.m["qpos0"])).repeat(nenv, 1)
        self.qvel = torch.zeros(nenv, cm.m["nv"], dtype=torch.float64)
        self.params = {}

    def set_param(self, name, rows):
        self.params[name] = rows.clone()

    def set_const(self, fields):
        out = {f: [] for f in fields}
        for e in range(self.nenv):
            m = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in self.model.host.items()}
            for n, rows in self.params.items():
                m[n] = rows[e].numpy().reshape(m[n].shape).copy()
            mjcf.set_const(m)
            for f in fields:
                out[f].append(np.asarray(m[f]).reshape(-1).copy())
        for f in fields:
            self.params[f] = self.torch.tensor(np.stack(out[f]))
        return {f: self.params[f] for f in fields}


HALVES = [[0.0254] * NOBJ, [0.02, 0.03, 0.0254, 0.035], [[0.03, 0.02, 0.025]] * NOBJ]
DENS = [1000.0, 700.0, 1500.0]


def _scene(nenv=3):
    from robogym_b200.rearrange_scene import BatchedBlockScene

    base = mjcf.compile_mjcf(scene_xml())
    sim = _RecordingSim(base, nenv)
    scene = BatchedBlockScene(sim)
    hs = np.zeros((nenv, NOBJ, 3))
    for e in range(nenv):
        for k in range(NOBJ):
            hs[e, k] = HALVES[e][k]
    scene.set_blocks(hs, density=np.array(DENS)[:, None])
    return base, sim, scene, hs


def test_rows_equal_the_separately_compiled_models():
    base, sim, scene, hs = _scene()
    assert scene.nobj == NOBJ
    for e in range(3):
        cm = mjcf.compile_mjcf(scene_xml(half=HALVES[e], density=DENS[e]))
        for name in ("geom_size", "geom_rbound", "geom_aabb", "body_mass", "body_inertia", "body_iquat", "body_subtreemass", "dof_invweight0", "body_invweight0", "opt_meaninertia"):
            got, want = sim.params[name][e].numpy(), np.asarray(cm.m[name], dtype=np.float64).reshape(-1)
            assert np.allclose(got, want, rtol=1e-10, atol=1e-14), (e, name, np.abs(got - want).max())
    # rescale_object_sizes: geometry only, masses as compiled
    mass_before = sim.params["body_mass"].clone()
    scene.rescale(np.array([[1.0, 1.2, 0.8, 1.0]] * 3))
    g1 = scene.geoms[1]
    assert np.allclose(sim.params["geom_size"][0].view(-1, 3)[g1].numpy(), 1.2 * 0.0254)
    assert np.allclose(sim.params["geom_rbound"][0][g1].item(), 1.2 * 0.0254 * np.sqrt(3))
    assert (sim.params["body_mass"] == mass_before).all()
    # materials
    scene.set_material(friction=np.array([[1.0, 0.005, 0.0001], [0.5, 0.01, 0.001], [2.0, 0.0, 0.0]]), solref=np.array([[0.02, 1.0], [-4000.0, -200.0], [0.01, 0.8]]))
    gf = sim.params["geom_friction"].view(3, -1, 3)
    assert np.allclose(gf[1, scene.geoms[2]].numpy(), [0.5, 0.01, 0.001]) and np.allclose(gf[1, 0].numpy(), np.asarray(base.m["geom_friction"]).reshape(-1, 3)[0])
    assert np.allclose(sim.params["geom_solref"].view(3, -1, 2)[1, scene.geoms[0]].numpy(), [-4000.0, -200.0])


def test_placement_parks_the_unused_blocks_on_the_floor():
    import torch

    base, sim, scene, hs = _scene()
    active = torch.tensor([[True, True, True, True], [True, True, False, False], [True, False, False, False]])
    xy = np.tile(np.array([[0.0, 0.0], [0.1, 0.0], [0.2, 0.0], [0.3, 0.0]]), (3, 1, 1))
    scene.place(xy, np.zeros((3, NOBJ)), TABLE_TOP + hs[:, :, 2] + 1e-3, active)
    for e in range(3):
        for k in range(NOBJ):
            p = sim.qpos[e, scene.qadr[k]:scene.qadr[k] + 3].numpy()
            if active[e, k]:
                assert np.allclose(p, [0.1 * k, 0.0, TABLE_TOP + hs[e, k, 2] + 1e-3])
            else:
                assert p[0] >= 3.0 and abs(p[2] - (hs[e, k, 2] + 1e-3)) < 1e-12     # resting height on the floor, far from the table


def _patched_emu(base_blob, base_m, rows, e):
    em = pyemu.EmuBatch(base_blob, base_m, 1, contact_capacity=64, row_capacity=160)
    for n, r in rows.items():
        em.model_field(n, np.float32)[:] = r[e].numpy().astype(np.float32)
    return em


def test_emulated_kernel_on_patched_rows_matches_the_oracle_on_compiled_models():
    """every environment of the padded batch behaves like the model the reference would have compiled for it"""
    import torch

    base, sim, scene, hs = _scene()
    scene.set_material(friction=np.array([[1.0, 0.005, 0.0001], [0.5, 0.01, 0.001], [2.0, 0.0, 0.0]]))
    fr = [(1.0, 0.005, 0.0001), (0.5, 0.01, 0.001), (2.0, 0.0, 0.0)]
    active = torch.tensor([[True, True, True, True], [True, True, True, False], [True, True, False, False]])
    xy = np.tile(np.array([[0.0, 0.0], [0.1, 0.05], [0.2, -0.05], [0.3, 0.0]]), (3, 1, 1))
    # dropped from 2 mm, the second block lands tilted on the first one's edge region of the table: contacts, friction, stacking
    scene.place(xy, np.array([[0.0, 0.4, 0.8, 1.2]] * 3), TABLE_TOP + hs[:, :, 2] + 2e-3, active)
    blob = base.blob()
    for e in range(3):
        cm = mjcf.compile_mjcf(scene_xml(half=HALVES[e], density=DENS[e], friction=fr[e]))
        om, d = oracle_pair(cm.blob())
        d.qpos[:] = sim.qpos[e].numpy()
        em = _patched_emu(blob, base.m, sim.params, e)
        em.qpos[0] = sim.qpos[e].numpy()
        # teacher-forced env-steps of 10 substeps
        worst = 0.0
        for k in range(12):
            em.qpos[0] = d.qpos; em.qvel[0] = d.qvel; em.warm[0] = d.qacc_warmstart
            for _ in range(10):
                d.step()
            d.forward()
            em.step(10, 1)
            worst = max(worst, np.abs(em.qpos[0] - d.qpos).max())
        assert em.warn.max() == 0 and d.warning[0] == 0
        assert worst < 2e-5, (e, worst)
        # the parked blocks sit on the floor and stay there
        for k in range(NOBJ):
            if not active[e, k]:
                assert abs(d.qpos[scene.qadr[k] + 2] - hs[e, k, 2]) < 2e-3 and d.qpos[scene.qadr[k]] > 2.9


@pytest.mark.gpu
def test_cuda_batch_with_a_different_scene_per_environment():
    import torch

    from robogym_b200 import build, engine
    from robogym_b200.rearrange_scene import BatchedBlockScene

    build.build()
    base = mjcf.compile_mjcf(scene_xml())
    nenv = 3
    model = engine.DeviceModel(base.blob(), 0)
    sim = engine.BatchedSim(model, nenv, 10, outputs=("ncon", "warn"), contact_capacity=64, row_capacity=160)
    scene = BatchedBlockScene(sim)
    hs = np.zeros((nenv, NOBJ, 3))
    for e in range(nenv):
        for k in range(NOBJ):
            hs[e, k] = HALVES[e][k]
    consts = scene.set_blocks(hs, density=np.array(DENS)[:, None])
    fr = [(1.0, 0.005, 0.0001), (0.5, 0.01, 0.001), (2.0, 0.0, 0.0)]
    scene.set_material(friction=np.array(fr))
    torch.cuda.synchronize()
    active = torch.tensor([[True, True, True, True], [True, True, True, False], [True, True, False, False]])
    xy = np.tile(np.array([[0.0, 0.0], [0.1, 0.05], [0.2, -0.05], [0.3, 0.0]]), (3, 1, 1))
    oracles = []
    for e in range(nenv):
        cm = mjcf.compile_mjcf(scene_xml(half=HALVES[e], density=DENS[e], friction=fr[e]))
        # mj_setConst on the device from the per-environment rows == the compiler's constants for that model
        for name in ("dof_invweight0", "body_invweight0", "body_subtreemass"):
            got, want = consts[name][e].cpu().numpy().astype(np.float64), np.asarray(cm.m[name]).reshape(-1)
            assert np.allclose(got, want, rtol=5e-4, atol=1e-6), (e, name, np.abs(got - want).max())
        oracles.append(oracle_pair(cm.blob()))
    scene.place(xy, np.array([[0.0, 0.4, 0.8, 1.2]] * 3), TABLE_TOP + hs[:, :, 2] + 2e-3, active)
    for e, (om, d) in enumerate(oracles):
        d.qpos[:] = sim.qpos[e].cpu().numpy().astype(np.float64)
    worst = np.zeros(nenv)
    for k in range(12):
        for e, (om, d) in enumerate(oracles):      # teacher-forced from the oracles
            sim.qpos[e] = torch.tensor(d.qpos, dtype=torch.float32); sim.qvel[e] = torch.tensor(d.qvel, dtype=torch.float32)
            sim.qacc_warmstart[e] = torch.tensor(d.qacc_warmstart, dtype=torch.float32)
        sim.step()
        torch.cuda.synchronize()
        q = sim.qpos.cpu().numpy().astype(np.float64)
        for e, (om, d) in enumerate(oracles):
            for _ in range(10):
                d.step()
            d.forward()
            worst[e] = max(worst[e], np.abs(q[e] - d.qpos).max())
    assert int(sim.warn.max()) == 0
    assert worst.max() < 5e-5, worst
