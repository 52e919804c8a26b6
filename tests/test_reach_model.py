"""dactyl/reach (BASELINE.json configs[0]: hand only, no cube) -- the second compiled model through the same
engines: kernel logic in CPU emulation and the CUDA path, both against the fp64 oracle, teacher-forced.
Without contacts between hand and cube the fp32/fp64 gap is round-off only, so the tolerances are tighter
than for dactyl/locked."""
import json
import os

import numpy as np
import pytest

import pyemu
from helpers import rollout_states
from robogym_b200 import modelblob

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def reach():
    blob = open(os.path.join(ROOT, "robogym_b200", "assets", "dactyl_reach.rgm"), "rb").read()
    names = json.load(open(os.path.join(ROOT, "robogym_b200", "assets", "dactyl_reach.names.json")))
    m = modelblob.unpack(blob)
    states, after, om = rollout_states(blob, 16, seed=3)
    return blob, names, m, states, after


def test_reach_model_dimensions(reach):
    blob, names, m, _, _ = reach
    assert (m["nq"], m["nv"], m["nu"]) == (24, 24, 20)          # ShadowHand alone: 24 joints, 20 actuators
    assert all(n.startswith("robot0:") for n in names["joint"])


def test_emulated_kernel_matches_oracle_on_reach(reach):
    blob, names, m, states, after = reach
    dims = {k: m[k] for k in modelblob.DIMS}
    e = pyemu.EmuBatch(blob, dims, len(states))
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st
    e.step(10, 1)
    eq = np.array([np.abs(e.qpos[k] - after[k][0]).max() for k in range(len(after))])
    ev = np.array([np.abs(e.qvel[k] - after[k][1]).max() for k in range(len(after))])
    assert int(e.warn.max()) == 0
    assert np.median(eq) < 2e-5 and eq.max() < 5e-4 and np.median(ev) < 1e-3


@pytest.mark.gpu
def test_cuda_matches_oracle_on_reach(reach):
    import torch

    from robogym_b200 import build, engine

    build.build()
    blob, names, m, states, after = reach
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, len(states), 10, outputs=("site_xpos", "ncon", "warn"))
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))
    sim.step()
    torch.cuda.synchronize()
    q, v = sim.qpos.cpu().numpy(), sim.qvel.cpu().numpy()
    eq = np.array([np.abs(q[k] - after[k][0]).max() for k in range(len(after))])
    ev = np.array([np.abs(v[k] - after[k][1]).max() for k in range(len(after))])
    assert int(sim.warn.max()) == 0
    assert np.median(eq) < 2e-5 and eq.max() < 5e-4 and np.median(ev) < 1e-3
    # 1000 random-action env-steps (the configuration's horizon) stay finite and inside the joint limits (+ soft margin)
    gen = torch.Generator(device=sim.device); gen.manual_seed(0)
    cr = torch.tensor(m["actuator_ctrlrange"].reshape(-1, 2), dtype=torch.float32, device=sim.device)
    for _ in range(1000):
        sim.ctrl.copy_(cr[:, 0] + (cr[:, 1] - cr[:, 0]) * torch.rand(len(states), 20, device=sim.device, generator=gen))
        sim.step()
    torch.cuda.synchronize()
    assert torch.isfinite(sim.qpos).all() and int(sim.warn.max()) == 0
    jr = torch.tensor(m["jnt_range"].reshape(-1, 2), dtype=torch.float32, device=sim.device)
    assert bool(((sim.qpos > jr[:, 0] - 0.2) & (sim.qpos < jr[:, 1] + 0.2)).all())
