"""CPU emulation of the CUDA device code (tests/emu) against the fp64 oracle: the kernel's fp32
logic is checked stage by stage and through teacher-forced env-steps before it ever reaches a GPU."""
import numpy as np
import pytest

import pyemu
from helpers import live_indices, oracle_pair, rollout_states, step_errors
from robogym_b200 import modelblob


@pytest.fixture(scope="module")
def setup(locked_blob):
    m = modelblob.unpack(locked_blob)
    dims = {k: m[k] for k in modelblob.DIMS}
    states, after, om = rollout_states(locked_blob, 24, seed=5)
    return m, dims, states, after, om


def load(e, k, st):
    e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st


def test_stagewise_parity(locked_blob, setup):
    m, dims, states, after, om = setup
    _, d = oracle_pair(locked_blob)
    e = pyemu.EmuBatch(locked_blob, dims, 1)
    for st in states[::6]:
        load(e, 0, st)
        d.qpos[:], d.qvel[:], d.ctrl[:] = st[0], st[1], st[2]
        d.userdata[:60] = st[3]
        d.qacc_warmstart[:] = st[4]
        d.forward()
        e.forward()
        g = e.dbg_view()
        rel = lambda a, b: np.abs(np.asarray(a, float) - b).max() / max(np.abs(b).max(), 1e-12)
        assert rel(e.site_xpos[0].ravel(), d.site_xpos) < 1e-6
        assert rel(g["M"].ravel(), d.M) < 1e-5
        assert rel(g["tlen"], d.ten_length) < 1e-6
        assert rel(g["tJ"].ravel(), d.ten_J) < 1e-5
        assert rel(g["bias"], d.qfrc_bias) < 1e-4
        assert rel(g["aforce"], d.actuator_force) < 1e-4
        assert rel(g["smooth"], d.qfrc_smooth) < 1e-4
        assert g["ncon"] == d.ncon[0]
        oc = d.contact.reshape(-1, 24)[:d.ncon[0]]
        for k in range(g["ncon"]):
            assert (int(g["con"][k, 20]), int(g["con"][k, 21])) == (int(oc[k, 20]), int(oc[k, 21]))
            assert abs(g["con"][k, 0] - oc[k, 0]) < 2e-6            # penetration distance (metres)
            assert np.abs(g["con"][k, 4:7] - oc[k, 4:7]).max() < 2e-3   # normal: MPR tolerance 1e-6 m over ~1e-2 m facets


def test_teacher_forced_env_step(locked_blob, locked_names, setup):
    """One env-step (10 x mj_step + forward) from identical states: fp32 kernel logic vs fp64 oracle."""
    m, dims, states, after, om = setup
    K = len(states)
    e = pyemu.EmuBatch(locked_blob, dims, K)
    for k, st in enumerate(states):
        load(e, k, st)
    e.step(10, 1)
    iq, iv = live_indices(om, locked_names)
    eq, ev = step_errors(e.qpos, e.qvel, after, iq, iv)
    assert e.warn.max() == 0
    assert np.median(eq) < 2e-4 and np.median(ev) < 5e-3
    assert np.mean(eq < 1e-3) > 0.7          # chaotic contact switching limits the tail, see DESIGN.md


def test_teacher_forced_contract_workload(locked_blob, locked_names, setup):
    """The GPU test's thresholds on the kernel logic in CPU emulation: 512 states from 8 seeds of the SURVEY 8(d) workload
    (full-range relative actions): >= 90 % within 1e-3 in qpos, >= 95 % with the oracle's contact count."""
    from helpers import contract_states

    m, dims, _, _, _ = setup
    sts, after, om = contract_states(locked_blob, range(100, 108), 64)
    e = pyemu.EmuBatch(locked_blob, dims, len(sts))
    for k, st in enumerate(sts):
        load(e, k, st)
    e.step(10, 1)
    iq, iv = live_indices(om, locked_names)
    eq, ev = step_errors(e.qpos, e.qvel, after, iq, iv)
    assert e.warn.max() == 0
    assert np.median(eq) < 2e-5 and np.median(ev) < 1e-3
    assert np.mean(eq < 1e-3) >= 0.90
    assert np.mean(e.ncon == np.array([a[2] for a in after])) >= 0.95


def test_result_is_independent_of_batch_slot(locked_blob, setup):
    """Determinism: the same state in different batch slots gives bitwise identical results."""
    m, dims, states, after, om = setup
    e = pyemu.EmuBatch(locked_blob, dims, 3)
    for k in range(3):
        load(e, k, states[4])
    load(e, 1, states[9])
    e.step(10, 1)
    assert np.array_equal(e.qpos[0], e.qpos[2]) and np.array_equal(e.qvel[0], e.qvel[2])
    assert not np.array_equal(e.qpos[0], e.qpos[1])


def test_bad_state_resets_and_flags(locked_blob, setup):
    m, dims, states, after, om = setup
    e = pyemu.EmuBatch(locked_blob, dims, 1)
    load(e, 0, states[0])
    e.qvel[0, 12] = np.nan
    e.step(2, 0)
    assert e.warn[0] & 4
    assert np.isfinite(e.qpos).all() and np.isfinite(e.qvel).all()


def test_external_force_and_per_env_timestep(locked_blob, locked_names, setup):
    """data.xfrc_applied (RandomizedWindWrapper pushes the cube with it) and a per-environment opt.timestep
    (RandomizedTimestepWrapper) -- kernel logic vs oracle, one env-step from identical states."""
    m, dims, states, after, om = setup
    cube = locked_names["body"].index("cube:middle")
    st = states[3]
    for xf, dt in (((0.3, -0.2, 0.5, 0.01, 0.0, -0.02), None), (None, 0.0065), ((0.0, 0.4, 0.0, 0.0, 0.03, 0.0), 0.0105)):
        om2, d = oracle_pair(locked_blob)
        d.qpos[:], d.qvel[:], d.ctrl[:] = st[0], st[1], st[2]
        d.userdata[:60] = st[3]
        d.qacc_warmstart[:] = st[4]
        e = pyemu.EmuBatch(locked_blob, dims, 1)
        load(e, 0, st)
        if xf is not None:
            d.xfrc_applied[6 * cube:6 * cube + 6] = xf
            e.xfrc = np.zeros((1, dims["nbody"], 6), np.float32)
            e.xfrc[0, cube] = xf
        if dt is not None:
            om2.field("opt_timestep")[0] = dt
            e.timestep = np.full(1, dt, np.float32)
        d.env_step(10)
        e.step(10, 1)
        iq, iv = live_indices(om, locked_names)
        assert np.abs(e.qpos[0][iq] - d.qpos[iq]).max() < 2e-3 and np.abs(e.qvel[0][iv] - d.qvel[iv]).max() < 8e-2
        # the push / the different step size must actually matter
        assert np.abs(d.qpos[iq] - after[3][0][iq]).max() > 1e-4
