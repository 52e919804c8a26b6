"""rg_set_const: mj_setConst per environment on the device (SimulationInterface.set_constants,
robogym/mujoco/simulation_interface.py:197-201; the reference calls it after its randomisers edited masses / inertias /
armatures, robogym/wrappers/randomizations.py).  Checked against the host restatement mjcf.set_const (numpy, fp64), which
is what fills these constants at compile time."""
import copy

import numpy as np
import pytest

import pyemu
from robogym_b200 import mjcf, modelblob

FIELDS = ("dof_invweight0", "body_invweight0", "tendon_invweight0", "tendon_length0", "body_subtreemass", "opt_meaninertia")


def _edited(m, rng):
    """what the randomisers touch: body masses and inertias (x0.5..2), armature, and the cube's size-dependent inertia"""
    m = copy.deepcopy(m)
    scale = rng.uniform(0.5, 2.0, m["nbody"])
    m["body_mass"] = m["body_mass"] * scale
    m["body_inertia"] = (m["body_inertia"].reshape(-1, 3) * (scale * rng.uniform(0.8, 1.25, m["nbody"]))[:, None]).reshape(-1)
    m["dof_armature"] = m["dof_armature"] * rng.uniform(0.5, 2.0, m["nv"])
    return m


def _want(m, names):
    ref = copy.deepcopy(m)
    cm = mjcf.CompiledModel.from_blob(modelblob.pack(ref, names), names)
    mjcf.set_const(cm.m, spatial_tendon_eval=lambda q: mjcf.tendon_eval(cm.m, q))
    return {k: np.asarray(cm.m[k], dtype=np.float64).reshape(-1) for k in FIELDS}


def _close(got, want, rtol):
    for k in FIELDS:
        if want[k].size == 0:
            continue
        err = np.abs(np.asarray(got[k], dtype=np.float64).reshape(-1) - want[k]) / np.maximum(np.abs(want[k]), 1e-12)
        assert err.max() < rtol, (k, float(err.max()))


def test_emulated_set_const_matches_the_host_restatement(locked_blob):
    names = modelblob.unpack_names(locked_blob)
    m0 = modelblob.unpack(locked_blob)
    # the committed blob's constants are mj_setConst of the unedited model
    e = pyemu.EmuBatch(locked_blob, {k: m0[k] for k in modelblob.DIMS}, 1)
    _close(e.set_const(), {k: np.asarray(m0[k], dtype=np.float64).reshape(-1) for k in FIELDS}, 2e-4)
    rng = np.random.RandomState(0)
    for _ in range(3):
        m = _edited(m0, rng)
        blob = modelblob.pack(m, names)
        e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, 1)
        got = e.set_const()
        _close(got, _want(m, names), 2e-4)
        assert np.abs(got["dof_invweight0"] / m0["dof_invweight0"] - 1).max() > 0.05      # the edit mattered


@pytest.mark.gpu
def test_cuda_set_const_per_environment_matches_the_host_restatement(locked_blob):
    import torch

    from robogym_b200 import build, engine

    build.build()
    names = modelblob.unpack_names(locked_blob)
    m0 = modelblob.unpack(locked_blob)
    n = 6
    rng = np.random.RandomState(1)
    edits = [_edited(m0, rng) for _ in range(n)]
    model = engine.DeviceModel(locked_blob, 0)
    sim = engine.BatchedSim(model, n, 10, outputs=("warn",))
    for name in ("body_mass", "body_inertia", "dof_armature"):
        sim.set_param(name, np.stack([np.asarray(e[name]).reshape(-1) for e in edits]))
    # only environments 1.. are recomputed; environment 0 keeps the shared model's constants
    mask = torch.ones(n, dtype=torch.uint8, device=sim.device)
    mask[0] = 0
    out = sim.set_const(mask=mask, fields=engine.BatchedSim.SET_CONST_FIELDS)
    torch.cuda.synchronize()
    for k in engine.BatchedSim.SET_CONST_FIELDS:
        if m0[k].size:
            assert np.allclose(out[k][0].cpu().numpy(), np.asarray(m0[k]).reshape(-1), rtol=1e-6), k
    for i in range(1, n):
        _close({k: out[k][i].cpu().numpy() for k in out}, _want(edits[i], names), 5e-4)
    # and the step consumes them: the launch runs with the recomputed rows (finite, no warning)
    sim.step()
    torch.cuda.synchronize()
    assert int(sim.warn.max()) == 0 and bool(torch.isfinite(sim.qpos).all())
