"""BASELINE.json configs[2] -- dactyl/full_perpendicular (Rubik's cube: 26 cubelets on 6 face drivers, nq170 / nv168,
~35 contacts with condim 6, ~500 constraint rows; robogym/envs/dactyl/full_perpendicular.py:92-154) -- on the engine.
The model blob is compiled from the reference's own env build (tools/compile_models.py); the engine runs it with run-time
capacities (rg_batch_create_ex: 64 contacts, 256 single-row elements, 32 dofs per contact) and one warp per SM
(111 KB of shared-memory scratch per environment).  Parity vs the fp64 oracle, teacher-forced."""
import os

import numpy as np
import pytest

from helpers import oracle_pair

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CAPS = dict(contact_capacity=64, row_capacity=256, dofs_per_contact=32)


@pytest.fixture(scope="module")
def full():
    from robogym_b200 import modelblob

    blob = open(os.path.join(ROOT, "robogym_b200", "assets", "dactyl_full_perpendicular.rgm"), "rb").read()
    m = modelblob.unpack(blob)
    om, d = oracle_pair(blob)
    nu = m["nu"]
    cr = m["actuator_ctrlrange"].reshape(-1, 2)
    rng = np.random.RandomState(3)
    d.ctrl[:] = cr.mean(1)
    for _ in range(5):
        d.env_step(10)
    states, after = [], []
    for _ in range(16):
        a = rng.uniform(-1, 1, nu)
        d.ctrl[:] = np.clip(d.ctrl + 0.3 * a * (cr[:, 1] - cr[:, 0]) / 2, cr[:, 0], cr[:, 1])
        states.append((d.qpos.copy(), d.qvel.copy(), d.ctrl.copy(), d.userdata[:3 * nu].copy(), d.qacc_warmstart.copy()))
        d.env_step(10)
        after.append((d.qpos.copy(), d.qvel.copy(), int(d.ncon[0])))
    return blob, m, states, after


def check(qpos, ncon, warn, after):
    eq = np.array([np.abs(qpos[k] - after[k][0]).max() for k in range(len(after))])
    dn = np.array([abs(int(ncon[k]) - after[k][2]) for k in range(len(after))])
    assert int(np.max(warn)) == 0                     # nothing overflowed
    assert (m_ := np.median(eq)) < 2e-3, m_           # one env-step of a 26-body pile, fp32 vs fp64
    assert eq.max() < 3e-2
    assert dn.max() <= 5 and np.median(dn) <= 2       # ~35 contacts, a few sit at the activation margin


def test_dimensions(full):
    blob, m, states, after = full
    assert (m["nq"], m["nv"], m["nu"], m["nbody"]) == (170, 168, 20, 135)
    assert min(a[2] for a in after) >= 10


def test_kernel_logic_matches_oracle_in_emulation(full):
    import pyemu
    from robogym_b200 import modelblob

    blob, m, states, after = full
    dims = {k: m[k] for k in modelblob.DIMS}
    e = pyemu.EmuBatch(blob, dims, len(states), **CAPS)
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st
    e.step(10, 1)
    check(e.qpos, e.ncon, e.warn, after)


def test_default_capacities_overflow_is_flagged_not_fatal(full):
    import pyemu
    from robogym_b200 import modelblob

    blob, m, states, after = full
    dims = {k: m[k] for k in modelblob.DIMS}
    e = pyemu.EmuBatch(blob, dims, 1)                 # 32 contacts / 64 rows / 16 dofs per contact
    e.qpos[0], e.qvel[0], e.ctrl[0], e.pid[0], e.warm[0] = states[0]
    e.step(10, 1)
    assert e.warn[0] & 3 and np.isfinite(e.qpos).all()


@pytest.mark.gpu
def test_cuda_matches_oracle(full):
    import torch

    from robogym_b200 import build, engine

    build.build()
    blob, m, states, after = full
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, len(states), 10, outputs=("site_xpos", "ncon", "warn"), **CAPS)
    assert sim.launch_info()["warps_per_cta"] >= 1
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))
    sim.step()
    torch.cuda.synchronize()
    check(sim.qpos.cpu().numpy(), sim.ncon.cpu().numpy(), sim.warn.cpu().numpy(), after)


def _randomised_rows(m, n, seed=5):
    """Per-environment parameter rows in the spirit of full_perpendicular.py:425-440 (the subset of that wrapper stack that
    is plain array scaling: robot / cube friction, gravity, robot damping, Kp, and a per-env timestep), drawn with the
    wrappers' distributions (randomizations.py:72-306: log-uniform factors, gravity +- 0.4 m/s^2 per axis)."""
    rng = np.random.RandomState(seed)
    rows = {}
    fr = np.tile(m["geom_friction"].reshape(1, -1, 3), (n, 1, 1))
    fr *= np.exp(rng.uniform(np.log(0.7), np.log(1.3), (n, 1, 1)))
    rows["geom_friction"] = fr.reshape(n, -1)
    rows["opt_gravity"] = m["opt_gravity"].reshape(1, 3) + rng.uniform(-0.4, 0.4, (n, 3))
    rows["dof_damping"] = m["dof_damping"].reshape(1, -1) * np.exp(rng.uniform(np.log(1 / 1.5), np.log(1.5), (n, m["nv"])))
    gp = np.tile(m["actuator_gainprm"].reshape(1, -1, 10), (n, 1, 1))
    gp[:, :, 0] *= np.exp(rng.uniform(np.log(0.5), np.log(2.0), (n, m["nu"])))
    rows["actuator_gainprm"] = gp.reshape(n, -1)
    ts = m["opt_timestep"][0] * rng.uniform(0.85, 1.15, n)
    return rows, ts


def _oracle_step_with_rows(blob, m, states, rows, ts, nsub=10):
    """One env-step per state on the oracle, each under its own parameter row."""
    after = []
    for k, st in enumerate(states):
        om, d = oracle_pair(blob)
        for name, v in rows.items():
            om.field(name)[:] = v[k]
        om.field("opt_timestep")[0] = ts[k]
        d.qpos[:], d.qvel[:], d.ctrl[:] = st[0], st[1], st[2]
        d.userdata[:3 * m["nu"]] = st[3]
        d.qacc_warmstart[:] = st[4]
        d.env_step(nsub)
        after.append((d.qpos.copy(), d.qvel.copy(), int(d.ncon[0])))
    return after


def test_kernel_logic_with_per_env_parameters_in_emulation(full):
    """The per-environment timestep path on the full cube (parameter ROWS exist only in the CUDA engine: the emulation has one
    model): each state stepped under its own opt.timestep."""
    import pyemu
    from robogym_b200 import modelblob

    blob, m, states, after, = full
    n = 6
    _, ts = _randomised_rows(m, n)
    want = _oracle_step_with_rows(blob, m, states[:n], {}, ts)
    dims = {k: m[k] for k in modelblob.DIMS}
    e = pyemu.EmuBatch(blob, dims, n, **CAPS)
    e.timestep = ts.astype(np.float32)
    for k, st in enumerate(states[:n]):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st
    e.step(10, 1)
    check(e.qpos, e.ncon, e.warn, want)


@pytest.mark.gpu
def test_cuda_matches_oracle_with_per_env_parameters(full):
    """cfg 3 with per-environment parameters (friction, gravity, damping, Kp rows through rg_batch_bind_param + a per-env
    timestep), every environment against an oracle that carries the same parameters."""
    import torch

    from robogym_b200 import build, engine

    build.build()
    blob, m, states, after = full
    n = len(states)
    rows, ts = _randomised_rows(m, n)
    want = _oracle_step_with_rows(blob, m, states, rows, ts)
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, n, 10, outputs=("site_xpos", "ncon", "warn"), **CAPS)
    for name, v in rows.items():
        sim.set_param(name, v)
    sim.enable_per_env_timestep().copy_(torch.tensor(ts, dtype=torch.float32, device=sim.device))
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))
    sim.step()
    torch.cuda.synchronize()
    check(sim.qpos.cpu().numpy(), sim.ncon.cpu().numpy(), sim.warn.cpu().numpy(), want)
    # and the parameters did matter: the same states without them land elsewhere
    base = np.array([a[0] for a in after])
    assert np.abs(np.array([w[0] for w in want]) - base).max() > 1e-3


@pytest.mark.gpu
def test_cuda_full_cube_at_batch_4096(full):
    """BASELINE.json configs[2] at its batch size (4096), per-environment parameters on: the 16 teacher-forcing states tiled over
    the batch must reproduce the small-batch result bit for bit in every slot (size-independent property), with no warning
    bits, and the slots that hold oracle states still agree with the oracle."""
    import torch

    from robogym_b200 import build, engine

    build.build()
    blob, m, states, after = full
    n, N = len(states), 4096
    rows, ts = _randomised_rows(m, n)
    want = _oracle_step_with_rows(blob, m, states, rows, ts)
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, N, 10, outputs=("site_xpos", "ncon", "warn"), **CAPS)
    idx = np.arange(N) % n
    for name, v in rows.items():
        sim.set_param(name, v[idx])
    sim.enable_per_env_timestep().copy_(torch.tensor(ts[idx], dtype=torch.float32, device=sim.device))
    f = lambda i: torch.tensor(np.stack([states[k][i] for k in idx]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))
    sim.step()
    torch.cuda.synchronize()
    q = sim.qpos.cpu().numpy()
    assert int(sim.warn.max()) == 0
    assert np.array_equal(q.reshape(N // n, n, -1), np.broadcast_to(q[:n], (N // n, n, q.shape[1])))
    check(q[:n], sim.ncon.cpu().numpy()[:n], sim.warn.cpu().numpy()[:n], want)
