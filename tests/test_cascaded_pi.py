"""mujoco-py's cascaded-PI user controller (actuator user="1"): the UR16e's default joint calibration
(robogym/assets/xmls/robot/ur16e/jointspec/calibrations/cascaded_pi/joint_actuations.xml:4-10, selected by
robogym/robot/robot_interface.py:61-63).  Its law lives in mujoco-py's mjpid.pyx, which is not in the reference tree; the
restatement here is PINNED by the reference's own impulse-response fixture (robogym/envs/rearrange/tests/test_rearrange_sim.py:135-230:
steady-state tool displacements 0.036 / 0.0363 / 0.022 / 0.022 +- 1e-3 and 90 % rise within 5 / 12 steps), which
tests/test_reference_suite.py runs on the shim with the reference's default calibration.  This file checks the law itself on a
small arm: closed-loop properties on the oracle, the kernel logic in CPU emulation against the oracle, and (gpu) CUDA against it."""
import numpy as np
import pytest

import pyemu
from helpers import oracle_pair
from robogym_b200 import mjcf, modelblob
from toy_models import CASCADED_ARM


@pytest.fixture(scope="module")
def arm():
    cm = mjcf.compile_mjcf(CASCADED_ARM)
    cm.m["opt_pid"][0] = 1          # what cymj.set_pid_control switches on
    return cm, cm.blob()


def _targets(k, rng):
    a = 0.11 * k
    return np.array([0.8 * np.sin(a), 0.5 + 0.4 * np.cos(0.7 * a), -0.6 * np.sin(1.3 * a), -0.02 - 0.02 * np.sin(a)]) + rng.uniform(-0.05, 0.05, 4) * [1, 1, 1, 0.1]


def rollout(blob, n, nsub=8):
    om, d = oracle_pair(blob)
    w = 6 * om.dim("nu")
    rng = np.random.RandomState(5)
    states, after = [], []
    for k in range(n):
        d.ctrl[:] = _targets(k, rng)
        states.append((d.qpos.copy(), d.qvel.copy(), d.ctrl.copy(), d.userdata[:w].copy(), d.qacc_warmstart.copy()))
        for _ in range(nsub):
            d.step()
        d.forward()
        after.append((d.qpos.copy(), d.qvel.copy(), d.userdata[:w].copy(), d.actuator_force.copy()))
    return states, after


def test_state_width_and_mixed_controllers(arm):
    cm, blob = arm
    assert modelblob.pid_stride(cm.m) == 6
    assert list(cm.m["actuator_user0"]) == [1, 1, 1, 0]
    om, _ = oracle_pair(blob)
    from oracle import pyoracle
    assert pyoracle.lib().rgo_pid_stride(om.ptr) == 6
    # a model without a cascaded actuator keeps mujoco-py's 3 floats per actuator
    plain = mjcf.compile_mjcf(CASCADED_ARM.replace(' user="1"', ""))
    assert modelblob.pid_stride(plain.m) == 3


def test_oracle_holds_the_arm_against_gravity_and_tracks_a_setpoint(arm):
    cm, blob = arm
    om, d = oracle_pair(blob)
    # zero set-point, arm stretched out horizontally: the bias-force term carries gravity from the first step on, so the light
    # joints (armature 0.01) do not sag while the velocity integrators are still empty
    d.forward()
    assert abs(d.actuator_force[1] + 0) > 1.0            # the shoulder motor pushes against a gravity torque of a few Nm
    assert abs(d.actuator_force[1] - d.qfrc_bias[1]) < 1e-9 and abs(d.actuator_force[2] - d.qfrc_bias[2]) < 1e-9
    for _ in range(200):
        d.step()
    assert np.abs(d.qpos[:3]).max() < 2e-3, d.qpos
    # a step in the set-point: the smoothed set-point (ema 0.97 per substep) reaches it, the velocity stays under max_vel, and the
    # joint settles on the target (position loop P-only on joints 1 and 3: no steady-state error without load because of the
    # bias-force term; joint 2 has integral action)
    d.ctrl[:3] = [1.0, -0.7, 0.9]
    vmax = np.zeros(3)
    for _ in range(3000):
        d.step()
        vmax = np.maximum(vmax, np.abs(d.qvel[:3]))
    assert np.all(vmax <= np.array([2.094, 3.142, 3.142]) * 1.2), vmax   # the velocity loop overshoots its clamped demand a little
    assert np.abs(d.qpos[:3] - [1.0, -0.7, 0.9]).max() < 5e-3, d.qpos
    ud = d.userdata[:24].reshape(4, 6)
    assert np.allclose(ud[:3, 4], [1.0, -0.7, 0.9], atol=1e-6)      # smoothed set-points
    assert np.all(ud[:, 5] == 1)                                      # "a step has been taken"
    # the gripper's plain PID shares the block (its 3 floats at the head of its 6)
    d.ctrl[3] = -0.03
    for _ in range(1500):
        d.step()
    assert abs(d.qpos[3] + 0.03) < 2e-3


def test_first_step_takes_the_setpoint_over_unsmoothed(arm):
    cm, blob = arm
    om, d = oracle_pair(blob)
    d.ctrl[:3] = [0.5, 0.4, -0.3]
    d.forward()                      # mj_forward runs the callback too; no step has been integrated yet
    assert np.allclose(d.userdata[:24].reshape(4, 6)[:3, 4], [0.5, 0.4, -0.3])
    d.step()
    assert np.allclose(d.userdata[:24].reshape(4, 6)[:3, 4], [0.5, 0.4, -0.3])
    d.ctrl[:3] = 0
    d.step()                         # from now on: ema * previous + (1 - ema) * ctrl
    assert np.allclose(d.userdata[:24].reshape(4, 6)[:3, 4], 0.97 * np.array([0.5, 0.4, -0.3]))
    d.reset()
    assert not d.userdata.any()


def test_emulated_kernel_matches_oracle_on_the_cascaded_arm(arm):
    cm, blob = arm
    states, after = rollout(blob, 80)
    e = pyemu.EmuBatch(blob, cm.m, len(states))
    assert e.pid.shape[1] == 24
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st
    e.step(8, 1)
    assert e.warn.max() == 0
    eq = np.abs(e.qpos - np.stack([a[0] for a in after])).max()
    ev = np.abs(e.qvel - np.stack([a[1] for a in after])).max()
    ep = np.abs(e.pid - np.stack([a[2] for a in after])).max()
    ef = np.abs(e.act_force - np.stack([a[3] for a in after]))
    assert eq < 2e-5 and ev < 5e-3 and ep < 5e-3, (eq, ev, ep)
    assert np.median(ef) < 1e-2 and ef.max() < 1.0, (np.median(ef), ef.max())


def test_emulated_free_run_tracks_the_oracle(arm):
    cm, blob = arm
    om, d = oracle_pair(blob)
    e = pyemu.EmuBatch(blob, cm.m, 1)
    e.qpos[0] = cm.m["qpos0"]
    rng = np.random.RandomState(5)
    for k in range(60):
        c = _targets(k, rng)
        d.ctrl[:] = c
        e.ctrl[0] = c
        for _ in range(8):
            d.step()
        d.forward()              # the callback runs (and advances its state) in mj_forward too, as in the reference's step + forward
        e.step(8, 1)
    assert np.abs(e.qpos[0] - d.qpos).max() < 5e-4


@pytest.mark.gpu
def test_cuda_matches_oracle_on_the_cascaded_arm(arm):
    import torch

    from robogym_b200 import build, engine

    build.build()
    cm, blob = arm
    states, after = rollout(blob, 80)
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, len(states), 8, outputs=("warn", "act_force"))
    assert sim.pid.shape[1] == 24
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))
    sim.step()
    torch.cuda.synchronize()
    assert int(sim.warn.max()) == 0
    eq = np.abs(sim.qpos.cpu().numpy() - np.stack([a[0] for a in after])).max()
    ev = np.abs(sim.qvel.cpu().numpy() - np.stack([a[1] for a in after])).max()
    ep = np.abs(sim.pid.cpu().numpy() - np.stack([a[2] for a in after])).max()
    assert eq < 5e-5 and ev < 1e-2 and ep < 1e-2, (eq, ev, ep)
    # reset clears the whole controller block (the flag included)
    sim.reset()
    torch.cuda.synchronize()
    assert not bool(sim.pid.any())
