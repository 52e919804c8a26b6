"""Shared test helpers: oracle rollouts that produce teacher-forcing states."""
import numpy as np

from oracle import pyoracle


def oracle_pair(blob):
    om = pyoracle.OracleModel(blob)
    return om, pyoracle.OracleData(om)


def rollout_states(blob, n_states, seed=1, settle=20, nsub=10, action_scale=0.3):
    """Roll the oracle with smoothed random position targets; return (states, after, model dims).

    states[k] = (qpos, qvel, ctrl, pid, warmstart) BEFORE env-step k, after[k] = (qpos, qvel, ncon) AFTER it."""
    om, d = oracle_pair(blob)
    nu = om.dim("nu")
    cr = om.field("actuator_ctrlrange").reshape(-1, 2)
    rng = np.random.RandomState(seed)
    d.ctrl[:] = cr.mean(1)
    for _ in range(settle):
        d.env_step(nsub)
    states, after = [], []
    for _ in range(n_states):
        a = rng.uniform(-1, 1, nu)
        d.ctrl[:] = np.clip(d.ctrl + action_scale * a * (cr[:, 1] - cr[:, 0]) / 2, cr[:, 0], cr[:, 1])
        states.append((d.qpos.copy(), d.qvel.copy(), d.ctrl.copy(), d.userdata[:3 * nu].copy(), d.qacc_warmstart.copy()))
        d.env_step(nsub)
        after.append((d.qpos.copy(), d.qvel.copy(), int(d.ncon[0])))
    return states, after, om


def live_indices(om, names):
    """qpos / qvel indices excluding the free-falling, collision-less target cube
    (robogym/envs/dactyl/locked.py:89-96; masked from observations, observation/mujoco.py:46,60)."""
    jq, jv = om.field("jnt_qposadr"), om.field("jnt_dofadr")
    jt = om.field("jnt_type")
    iq, iv = [], []
    for j, name in enumerate(names["joint"]):
        if name.startswith("target:"):
            continue
        nq, nv = {0: (7, 6), 1: (4, 3), 2: (1, 1), 3: (1, 1)}[int(jt[j])]
        iq += list(range(jq[j], jq[j] + nq))
        iv += list(range(jv[j], jv[j] + nv))
    return iq, iv


def step_errors(q_new, v_new, after, iq, iv):
    eq = np.array([np.abs(q_new[k][iq] - after[k][0][iq]).max() for k in range(len(after))])
    ev = np.array([np.abs(v_new[k][iv] - after[k][1][iv]).max() for k in range(len(after))])
    return eq, ev
