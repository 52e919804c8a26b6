"""Parity tests proper (run on the B200 box with -m gpu): the CUDA path, called through the C ABI
(robogym_b200.engine -> librobogym_b200.so), against the fp64 oracle on identical seeded states,
plus size-independent properties at BASELINE.json's full batch (8192)."""
import numpy as np
import pytest

from helpers import contract_states, live_indices, oracle_pair, rollout_states, step_errors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(locked_blob):
    import torch

    from robogym_b200 import build, engine

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    build.build()
    model = engine.DeviceModel(locked_blob, 0)
    return torch, engine, model


@pytest.fixture(scope="module")
def states(locked_blob):
    return rollout_states(locked_blob, 48, seed=11)


def put(torch, sim, sts):
    f = lambda i: torch.tensor(np.stack([s[i] for s in sts]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))


def test_native_library_is_loaded(gpu):
    torch, engine, model = gpu
    import ctypes

    assert isinstance(engine.lib(), ctypes.CDLL) and engine.LIB_PATH.endswith(".so")
    maps = open("/proc/self/maps").read()
    assert "librobogym_b200" in maps


def test_teacher_forced_env_step_vs_oracle(gpu, locked_blob, locked_names):
    """10 x mj_step + forward from identical states: 512 states from 8 seeds of the SURVEY 8(d) workload (full-range
    relative actions).  Tolerances (fp32 vs fp64, contact-rich): median |dq| < 2e-4 (rad / m), median |dv| < 5e-3,
    >= 90 % of the states within 1e-3 in qpos, >= 95 % with the oracle's contact count."""
    torch, engine, model = gpu
    sts, after, om = contract_states(locked_blob, range(200, 208), 64)
    sim = engine.BatchedSim(model, len(sts), 10, outputs=("site_xpos", "ncon", "warn"))
    put(torch, sim, sts)
    sim.step()
    torch.cuda.synchronize()
    iq, iv = live_indices(om, locked_names)
    eq, ev = step_errors(sim.qpos.cpu().numpy(), sim.qvel.cpu().numpy(), after, iq, iv)
    assert int(sim.warn.max()) == 0
    assert np.median(eq) < 2e-4 and np.median(ev) < 5e-3
    assert np.mean(eq < 1e-3) >= 0.90, np.mean(eq < 1e-3)
    ncon = sim.ncon.cpu().numpy()
    assert np.mean(ncon == np.array([a[2] for a in after])) >= 0.95


def test_stagewise_forward_vs_oracle(gpu, states, locked_blob):
    torch, engine, model = gpu
    sts, after, om = states
    sub = sts[::8]
    sim = engine.BatchedSim(model, len(sub), 10, outputs=("site_xpos", "act_force", "ncon", "warn"), debug=True)
    put(torch, sim, sub)
    sim.forward()
    torch.cuda.synchronize()
    _, d = oracle_pair(locked_blob)
    rel = lambda a, b: np.abs(np.asarray(a, float) - b).max() / max(np.abs(b).max(), 1e-12)
    for k, st in enumerate(sub):
        d.qpos[:], d.qvel[:], d.ctrl[:] = st[0], st[1], st[2]
        d.userdata[:60] = st[3]
        d.qacc_warmstart[:] = st[4]
        d.forward()
        g = sim.dbg_view(k)
        assert rel(sim.site_xpos[k].cpu().numpy().ravel(), d.site_xpos) < 1e-6
        assert rel(g["M"].ravel(), d.M) < 1e-5
        assert rel(g["tlen"], d.ten_length) < 1e-6
        assert rel(g["bias"], d.qfrc_bias) < 1e-4
        assert rel(g["smooth"], d.qfrc_smooth) < 1e-4
        assert rel(sim.act_force[k].cpu().numpy(), d.actuator_force) < 1e-4
        assert g["ncon"] == d.ncon[0]


def test_free_running_hand_only(gpu, locked_blob, locked_names):
    """T2 of SURVEY 8(d): 200 substeps free-running with the cube parked away from the hand
    (no contact-mode switches): hand qpos within 1e-3 of the oracle."""
    torch, engine, model = gpu
    om, d = oracle_pair(locked_blob)
    cr = om.field("actuator_ctrlrange").reshape(-1, 2)
    sim = engine.BatchedSim(model, 4, 10, outputs=("warn",))
    d.qpos[0] += 0.5
    sim.qpos[:, 0] += 0.5
    rng = np.random.RandomState(3)
    for _ in range(20):
        c = cr[:, 0] + (cr[:, 1] - cr[:, 0]) * rng.uniform(0.3, 0.7, len(cr))
        d.ctrl[:] = c
        sim.ctrl.copy_(torch.tensor(c, dtype=torch.float32, device=sim.device).repeat(4, 1))
        d.env_step(10)
        sim.step()
    torch.cuda.synchronize()
    q = sim.qpos.cpu().numpy()
    hand = [om.field("jnt_qposadr")[j] for j, n in enumerate(locked_names["joint"]) if n.startswith("robot0:")]
    assert np.abs(q[0][hand] - d.qpos[hand]).max() < 1e-3
    assert np.array_equal(q[0], q[3])


def test_full_batch_properties(gpu, states):
    """BASELINE.json config[1] size (8192 envs): identical inputs -> bitwise identical outputs in every
    slot (determinism / slot independence), no warnings, finite state, cubes stay on the palm."""
    torch, engine, model = gpu
    sts, after, om = states
    N = 8192
    sim = engine.BatchedSim(model, N, 10, outputs=("site_xpos", "ncon", "warn"))
    small = engine.BatchedSim(model, len(sts), 10, outputs=("warn",))
    put(torch, small, sts)
    idx = torch.arange(N, device=sim.device) % len(sts)
    for name in ("qpos", "qvel", "ctrl", "pid", "qacc_warmstart"):
        getattr(sim, name).copy_(getattr(small, name)[idx])
    for _ in range(3):
        sim.step()
        small.step()
    torch.cuda.synchronize()
    assert torch.equal(sim.qpos, small.qpos[idx]) and torch.equal(sim.qvel, small.qvel[idx])
    assert torch.isfinite(sim.qpos).all() and torch.isfinite(sim.qvel).all()
    assert int((sim.warn & 4).max()) == 0


def test_reset_restores_qpos0(gpu):
    torch, engine, model = gpu
    sim = engine.BatchedSim(model, 16, 10, outputs=("warn",))
    q0 = sim.qpos.clone()
    sim.ctrl.fill_(0.1)
    sim.step()
    mask = torch.zeros(16, dtype=torch.uint8, device=sim.device)
    mask[::2] = 1
    sim.reset(mask)
    torch.cuda.synchronize()
    assert torch.equal(sim.qpos[::2], q0[::2]) and not torch.equal(sim.qpos[1::2], q0[1::2])
    assert float(sim.qvel[::2].abs().max()) == 0.0 and float(sim.pid[::2].abs().max()) == 0.0


def test_model_edit_changes_dynamics(gpu):
    """Randomisers write model arrays in place (robogym/wrappers/randomizations.py:188): gravity here."""
    torch, engine, model = gpu
    sim = engine.BatchedSim(model, 2, 10, outputs=("warn",))
    sim.step()
    torch.cuda.synchronize()
    z1 = float(sim.qpos[0, 9])      # free-falling target cube z
    g0 = model.host["opt_gravity"].copy()
    try:
        model.set_field("opt_gravity", [0.0, 0.0, -1.0])
        sim2 = engine.BatchedSim(model, 2, 10, outputs=("warn",))
        sim2.step()
        torch.cuda.synchronize()
        z2 = float(sim2.qpos[0, 9])
    finally:
        model.set_field("opt_gravity", g0)
    assert z1 < 0 and z2 < 0 and abs(z1 / z2 - 9.81) < 1e-3


def test_per_env_parameter_overrides(gpu, locked_blob):
    """Domain randomisation: per-environment model arrays.  Gravity is checked in closed form on the
    ballistic target cube, joint damping against the oracle run with the same edited model."""
    torch, engine, model = gpu
    sim = engine.BatchedSim(model, 3, 10, outputs=("warn",))
    g = np.tile(model.host["opt_gravity"], (3, 1))
    g[1] = [0, 0, -1.0]
    g[2] = [0, 0, -20.0]
    sim.set_param("opt_gravity", g)
    damp = np.tile(model.host["dof_damping"], (3, 1))
    damp[2] *= 3.0
    sim.set_param("dof_damping", damp)
    cr = model.host["actuator_ctrlrange"].reshape(-1, 2)
    ctrl = cr[:, 0] + 0.7 * (cr[:, 1] - cr[:, 0])
    sim.ctrl.copy_(torch.tensor(ctrl, dtype=torch.float32, device=sim.device).repeat(3, 1))
    sim.qpos[:, 0] += 0.5          # park the cube: compare smooth hand dynamics only
    sim.step()
    torch.cuda.synchronize()
    h, n = 0.008, 10
    z = sim.qpos[:, 9].cpu().numpy()
    for k, gz in enumerate((-9.81, -1.0, -20.0)):
        assert abs(z[k] - gz * h * h * n * (n + 1) / 2) < 2e-5 * abs(gz)
    q = sim.qpos.cpu().numpy()
    hand = list(range(14, 38))
    for k in (0, 2):
        om, d = oracle_pair(locked_blob)
        om.field("opt_gravity")[:] = g[k]
        om.field("dof_damping")[:] = damp[k]
        d.ctrl[:] = ctrl
        d.qpos[0] += 0.5
        d.env_step(10)
        assert np.abs(q[k][hand] - d.qpos[hand]).max() < 2e-4
    assert np.abs(q[0][hand] - q[2][hand]).max() > 1e-3      # the override really changed the dynamics


def test_subset_launch_matches_full_launch(gpu, states):
    """rg_step_subset: selected environments get bit-identical results to a full launch, the others are untouched;
    an empty selection is a no-op; extra forwards (final_forward > 1) advance only the PID state and the solve."""
    torch, engine, model = gpu
    sts, _, _ = states
    n = len(sts)
    full = engine.BatchedSim(model, n, 10, outputs=("site_xpos", "act_force", "ncon", "warn"))
    part = engine.BatchedSim(model, n, 10, outputs=("site_xpos", "act_force", "ncon", "warn"))
    put(torch, full, sts); put(torch, part, sts)
    mask = torch.zeros(n, dtype=torch.bool, device=full.device)
    mask[::3] = True
    before = {k: getattr(part, k).clone() for k in ("qpos", "qvel", "pid", "qacc_warmstart")}
    full.step(final_forward=3)
    part.step(final_forward=3, mask=mask)
    torch.cuda.synchronize()
    for k in ("qpos", "qvel", "pid", "qacc_warmstart", "site_xpos", "act_force"):
        assert torch.equal(getattr(part, k)[mask], getattr(full, k)[mask]), k
    for k, v in before.items():
        assert torch.equal(getattr(part, k)[~mask], v[~mask]), k
    snap = {k: getattr(part, k).clone() for k in before}
    part.step(mask=torch.zeros(n, dtype=torch.bool, device=full.device))
    torch.cuda.synchronize()
    for k, v in snap.items():
        assert torch.equal(getattr(part, k), v), k
    # forwards leave qpos/qvel alone but advance the PID state
    q = part.qpos.clone(); p = part.pid.clone()
    part.forward(mask=mask, count=2)
    torch.cuda.synchronize()
    assert torch.equal(part.qpos, q) and not torch.equal(part.pid[mask], p[mask]) and torch.equal(part.pid[~mask], p[~mask])
