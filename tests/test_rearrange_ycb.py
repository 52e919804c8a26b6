"""BASELINE.json configs[4] -- the rearrange/ycb scene: the rearrange world of tests/test_rearrange_blocks.py with eight YCB mesh
objects (cracker box, banana, mug, power drill, hammer, soup can, scissors, apple: unions of 1..29 convex parts, 59 mesh geoms
in all) composed like MeshRearrangeSim.make_objects_xml / make_mesh_object do (tools/compose_reference_xml.py: rearrange_ycb_xml;
robogym/envs/rearrange/simulation/mesh.py:49-64, robogym/envs/rearrange/common/utils.py:250-281) and compiled to the committed
blob robogym_b200/assets/rearrange_ycb8.rgm: nq64 / nv56 as SURVEY 8a row a13 lists it.  Mesh-on-box resting contacts (one MPR
point per convex part), multi-geom free bodies, 163 geoms / 2888 candidate pairs."""
import os

import numpy as np
import pytest

import pyemu
from helpers import oracle_pair
from robogym_b200 import modelblob

HERE = os.path.dirname(os.path.abspath(__file__))
ASSET = os.path.join(HERE, "..", "robogym_b200", "assets", "rearrange_ycb8.rgm")
ARM_INIT = np.deg2rad(np.array([135.0, -90.0, 135.0, -100.0, -240.0, 135.0]))   # robogym/robot/ur16e/arm_interface.py:27
TABLE_TOP = 0.453 + 0.03324


@pytest.fixture(scope="module")
def scene():
    blob = open(ASSET, "rb").read()
    return blob, modelblob.unpack(blob), modelblob.unpack_names(blob)


def _adr(m, names, i):
    return int(m["jnt_qposadr"][names["joint"].index(f"object{i}:joint")])


def _reset(om, d, m, names):
    d.reset()
    d.qpos[:6] = ARM_INIT
    for i in range(8):                                   # out of the way while the tool pose is read
        d.qpos[_adr(m, names, i):_adr(m, names, i) + 3] = [1.0 + 0.25 * (i % 4), 1.1 + 0.3 * (i // 4), 0.8]
    d.forward()
    tcp = names["body"].index("robot0:gripper_tcp")
    om.field("eq_data")[:7] = [0, 0, 0, 1, 0, 0, 0]      # gym reset_mocap_welds / reset_mocap2body_xpos (mocap_solver.py:55-57)
    d.mocap_pos[:3] = d.xpos[3 * tcp:3 * tcp + 3]
    d.mocap_quat[:4] = d.xquat[4 * tcp:4 * tcp + 4]
    for i in range(8):
        b = names["body"].index(f"object{i}")
        zmin = 0.0
        for g in range(m["ngeom"]):
            if m["geom_bodyid"][g] == b:
                a, n = m["mesh_vertadr"][m["geom_dataid"][g]], m["mesh_vertnum"][m["geom_dataid"][g]]
                zmin = min(zmin, (m["mesh_vert"].reshape(-1, 3)[a:a + n, 2] + m["geom_pos"].reshape(-1, 3)[g, 2]).min())
        k = i if i < 4 else i + 1
        d.qpos[_adr(m, names, i):_adr(m, names, i) + 3] = [1.25 + 0.27 * (k % 3), 0.32 + 0.36 * (k // 3), TABLE_TOP - zmin + 0.002]
    d.warning[:] = 0
    return tcp


def test_scene_is_the_surveyed_model(scene):
    blob, m, names = scene
    assert (m["nq"], m["nv"], m["nu"]) == (64, 56, 1)                    # ycb, 8 objects
    per_object = [sum(1 for g in range(m["ngeom"]) if names["body"][m["geom_bodyid"][g]] == f"object{i}") for i in range(8)]
    assert per_object == [1, 3, 29, 5, 6, 1, 13, 1] and m["ngeom"] == 163 and m["npair"] > 2000
    assert m["opt_cone"][0] == 1 and m["nmocap"] == 1 and m["neq"] == 2
    # the combined centre of mass of every object sits at its body origin (make_mesh_object shifts the geoms by -center_mass)
    ipos = m["body_ipos"].reshape(-1, 3)
    for i in range(8):
        assert np.abs(ipos[names["body"].index(f"object{i}")]).max() < 2e-3, i


def _rollout(blob, m, names, n, settle=400):
    om, d = oracle_pair(blob)
    _reset(om, d, m, names)
    eqd = om.field("eq_data").copy()
    for _ in range(settle):
        d.step()
    p0 = d.mocap_pos[:3].copy()
    lo, hi = m["actuator_ctrlrange"].reshape(-1, 2)[0]
    rng = np.random.RandomState(0)
    states, after = [], []
    for k in range(n):
        a = 0.15 * k
        d.mocap_pos[:3] = p0 + [0.03 * np.sin(a), 0.04 * (1 - np.cos(a)), -0.03 * np.sin(0.5 * a)]
        d.ctrl[0] = rng.uniform(lo, hi)
        states.append((d.qpos.copy(), d.qvel.copy(), d.ctrl.copy(), d.userdata[:3].copy(), d.qacc_warmstart.copy(),
                       d.mocap_pos[:3].copy(), d.mocap_quat[:4].copy()))
        for _ in range(20):
            d.step()
        d.forward()
        after.append((d.qpos.copy(), d.qvel.copy(), int(d.ncon[0])))
    assert d.warning[0] == 0
    return states, after, eqd, d


def _errors(q, after, m, names):
    arm = np.array([np.abs(q[k][:8] - after[k][0][:8]).max() for k in range(len(after))])
    obj = np.array([max(np.abs(q[k][_adr(m, names, i):_adr(m, names, i) + 3] - after[k][0][_adr(m, names, i):_adr(m, names, i) + 3]).max()
                        for i in range(8)) for k in range(len(after))])
    return arm, obj


def test_oracle_objects_rest_on_the_table(scene):
    blob, m, names = scene
    _, after, _, d = _rollout(blob, m, names, 4, settle=1200)
    for i in range(8):
        z = d.qpos[_adr(m, names, i) + 2]
        assert TABLE_TOP < z < TABLE_TOP + 0.12, (i, z)
    assert np.abs(d.qpos[:6] - ARM_INIT).max() < 0.1            # the arm followed the few centimetres the mocap target moved
    assert 15 <= after[-1][2] <= 40                              # one point per resting convex part + the gripper pads


def test_emulated_kernel_matches_oracle_on_the_ycb_scene(scene):
    blob, m, names = scene
    states, after, eqd, _ = _rollout(blob, m, names, 16)
    e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, len(states), contact_capacity=64, row_capacity=128)
    e.model_field("eq_data", np.float32)[:] = eqd
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st[:5]
        e.mocap_pos[k, 0], e.mocap_quat[k, 0] = st[5], st[6]
    e.step(20, 1)
    arm, obj = _errors(e.qpos, after, m, names)
    assert e.warn.max() == 0
    assert arm.max() < 2e-5 and np.median(obj) < 5e-5 and obj.max() < 1e-3, (arm.max(), np.median(obj), obj.max())
    assert np.mean(np.abs(e.ncon - np.array([a[2] for a in after])) <= 2) > 0.8


@pytest.mark.gpu
def test_cuda_matches_oracle_on_the_ycb_scene(scene):
    import torch

    from robogym_b200 import build, engine

    build.build()
    blob, m, names = scene
    states, after, eqd, _ = _rollout(blob, m, names, 32)
    model = engine.DeviceModel(blob, 0)
    model.set_field("eq_data", eqd)
    sim = engine.BatchedSim(model, len(states), 20, outputs=("ncon", "warn"), contact_capacity=64, row_capacity=128)
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))
    sim.mocap_pos[:, 0].copy_(f(5)); sim.mocap_quat[:, 0].copy_(f(6))
    sim.step()
    torch.cuda.synchronize()
    arm, obj = _errors(sim.qpos.cpu().numpy(), after, m, names)
    assert int(sim.warn.max()) == 0
    assert arm.max() < 1e-4 and np.median(obj) < 1e-4 and np.mean(obj < 1e-3) > 0.9, (arm.max(), np.median(obj), obj.max())
    assert np.mean(np.abs(sim.ncon.cpu().numpy() - np.array([a[2] for a in after])) <= 2) > 0.8
