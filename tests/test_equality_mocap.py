"""Equality constraints and mocap bodies -- the physics features the rearrange scenes add to the path
(weld + joint equality, mocap body; robogym/assets/xmls/robot/ur16e/base.xml:52-54,
gripper_actuators.xml:2-4, tcp_mocap.xml:2, robogym/robot/control/tcp/mocap_solver.py:41-56) -- on a small inline model:
closed forms on the oracle, the kernel logic in CPU emulation against the oracle, and (gpu) the CUDA engine against it."""
import numpy as np
import pytest

import pyemu
from helpers import oracle_pair
from robogym_b200 import mjcf, modelblob
from toy_models import MOCAP_ARM


@pytest.fixture(scope="module")
def arm():
    cm = mjcf.compile_mjcf(MOCAP_ARM)
    return cm, cm.blob()


def _target(k):
    """mocap trajectory: the tool target circles and tilts, the hook carries the brick up and sideways"""
    a = 0.05 * k
    pos = np.array([[0.40 + 0.05 * np.cos(a), 0.08 * np.sin(a), 0.50 + 0.04 * np.sin(0.7 * a)],
                    [-0.3 + 0.002 * k, 0.2, 0.6 + 0.001 * k]])
    half = 0.15 * np.sin(0.5 * a)
    quat = np.array([[np.cos(half), 0.0, np.sin(half), 0.0], [1.0, 0.0, 0.0, 0.0]])
    return pos, quat


def rollout(blob, n, nsub=5):
    om, d = oracle_pair(blob)
    states, after = [], []
    rng = np.random.RandomState(3)
    for k in range(n):
        pos, quat = _target(k)
        d.mocap_pos[:] = pos.ravel()
        d.mocap_quat[:] = quat.ravel()
        d.ctrl[:] = rng.uniform(0, 0.02, om.dim("nu"))
        states.append((d.qpos.copy(), d.qvel.copy(), d.ctrl.copy(), np.zeros(3 * om.dim("nu")), d.qacc_warmstart.copy(), pos.copy(), quat.copy()))
        for _ in range(nsub):
            d.step()
        d.forward()
        after.append((d.qpos.copy(), d.qvel.copy(), int(d.nefc[0])))
    return states, after, om, d


def test_compiler_reads_welds_couplings_and_mocap_bodies(arm):
    cm, _ = arm
    m = cm.m
    assert m["nmocap"] == 2 and m["neq"] == 3
    assert list(m["eq_type"]) == [mjcf.EQ_WELD, mjcf.EQ_WELD, mjcf.EQ_JOINT]
    assert m["body_mocapid"][cm.name2id("body", "mocap")] == 0 and m["body_mocapid"][cm.name2id("body", "hook")] == 1
    # relative pose of body 2 in the frame of body 1 at qpos0 (mj_setConst): the tcp sits at (0.45, 0, 0.5), the mocap body 5 cm above
    w = m["eq_data"].reshape(-1, 7)
    assert np.allclose(w[0], [0, 0, -0.05, 1, 0, 0, 0], atol=1e-12)
    assert np.allclose(w[1], [0, 0, 0, 1, 0, 0, 0], atol=1e-12)
    assert np.allclose(w[2, :5], [0, 1, 0, 0, 0])


def _weld_sag(mass_invweight, g=9.81, solref=(0.02, 1.0), solimp=(0.9, 0.95, 0.001, 0.5, 2.0)):
    """A free body hanging from a fixed mocap body by a weld: at rest the constraint carries m g, force = aref / R with
    aref = -k d r and R = (1 - d) / d * A, A = body_invweight0 = 1/m for a free body, so r = g (1 - d) / (k d^2), d = d(|r|)."""
    d0, d1, width, mid, power = solimp
    k = 1.0 / (d1 ** 2 * solref[0] ** 2 * solref[1] ** 2)
    r = 1e-4
    for _ in range(200):
        x = min(r / width, 1.0)
        y = x ** power / mid ** (power - 1) if x <= mid else 1.0 - (1.0 - x) ** power / (1.0 - mid) ** (power - 1)
        d = d0 + y * (d1 - d0)
        r = g * (1.0 - d) / (k * d * d)
    return r


def test_oracle_weld_sags_by_the_closed_form_and_the_tool_follows_the_mocap(arm):
    cm, blob = arm
    om, d = oracle_pair(blob)
    brick, tcp = cm.name2id("body", "brick"), cm.name2id("body", "tcp")
    for _ in range(3000):
        d.step()
    d.forward()
    assert d.warning[0] == 0 and np.abs(d.qvel).max() < 1e-7
    sag = d.mocap_pos[5] - d.xpos[3 * brick + 2]
    assert abs(sag - _weld_sag(None)) < 1e-8, (sag, _weld_sag(None))
    # the tool: weld pulls the tcp to 5 cm below the mocap body, against gravity on a compliant arm -> millimetres
    want = d.mocap_pos[:3] + np.array([0, 0, -0.05])
    assert np.abs(d.xpos[3 * tcp:3 * tcp + 3] - want).max() < 5e-3
    # now ask for a pose the 4-joint arm can reach exactly (taken from its own forward kinematics): the tool goes there,
    # position and orientation, up to the millimetres gravity costs against the soft weld
    from robogym_b200.mjcf import rot_vec
    _, d2 = oracle_pair(blob)
    d2.qpos[:4] = [0.3, -0.4, 0.6, 0.2]
    d2.forward()
    qt, xt = d2.xquat[4 * tcp:4 * tcp + 4].copy(), d2.xpos[3 * tcp:3 * tcp + 3].copy()
    d.mocap_quat[:4] = qt
    d.mocap_pos[:3] = xt - rot_vec(qt, np.array([0, 0, -0.05]))
    for _ in range(4000):
        d.step()
    d.forward()
    assert np.abs(d.qvel).max() < 1e-6
    assert abs(abs(d.xquat[4 * tcp:4 * tcp + 4] @ qt) - 1.0) < 1e-4          # same orientation (up to sign)
    assert np.abs(d.xpos[3 * tcp:3 * tcp + 3] - xt).max() < 5e-3
    assert np.abs(d.qpos[:4] - [0.3, -0.4, 0.6, 0.2]).max() < 0.05


def test_oracle_joint_coupling_makes_the_second_finger_mirror_the_first(arm):
    cm, blob = arm
    om, d = oracle_pair(blob)
    r, l = cm.m["jnt_qposadr"][cm.name2id("joint", "r_slide")], cm.m["jnt_qposadr"][cm.name2id("joint", "l_slide")]
    d.ctrl[0] = 0.015
    for _ in range(2000):
        d.step()
    assert d.qpos[r] > 0.01 and abs(d.qpos[r] - d.qpos[l]) < 2e-4, (d.qpos[r], d.qpos[l])
    # without the coupling the left finger has no reason to move
    om.field("eq_active")[2] = 0
    d.reset()
    d.ctrl[0] = 0.015
    for _ in range(500):
        d.step()
    assert d.qpos[r] > 0.01 and abs(d.qpos[l]) < 1e-3


def test_emulated_kernel_matches_oracle_on_the_mocap_arm(arm):
    cm, blob = arm
    states, after, om, _ = rollout(blob, 60)
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, len(states))
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k], e.mocap_pos[k], e.mocap_quat[k] = st
    e.step(5, 1)
    eq = np.array([np.abs(e.qpos[k] - after[k][0]).max() for k in range(len(states))])
    ev = np.array([np.abs(e.qvel[k] - after[k][1]).max() for k in range(len(states))])
    assert e.warn.max() == 0
    assert eq.max() < 2e-5 and ev.max() < 5e-3, (eq.max(), ev.max())
    # 2 welds x 6 rows + 1 coupling, each as a pair of one-sided rows, + shoulder / finger limits when they are near
    nel = [e.dbg_view(k)["nel"] for k in range(len(states))]
    assert min(nel) >= 26


def test_emulated_free_run_tracks_the_oracle_and_the_weld_can_be_switched_off(arm):
    cm, blob = arm
    n = 40
    om, d = oracle_pair(blob)
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, 1)
    e.qpos[0] = cm.m["qpos0"]
    for k in range(n):
        pos, quat = _target(k)
        d.mocap_pos[:] = pos.ravel(); d.mocap_quat[:] = quat.ravel()
        e.mocap_pos[0] = pos; e.mocap_quat[0] = quat
        for _ in range(5):
            d.step()
        e.step(5, 1)
    assert np.abs(e.qpos[0] - d.qpos).max() < 2e-4
    # eq_active = 0 on the hook weld (robogym/envs/rearrange/common/base.py:452 does this to the model): the brick falls
    e.model_field("eq_active", np.int32)[1] = 0
    z0 = float(e.qpos[0, cm.m["nq"] - 7 + 2])
    e.step(50, 1)
    assert float(e.qpos[0, cm.m["nq"] - 7 + 2]) < z0 - 0.03


@pytest.mark.gpu
def test_cuda_matches_oracle_on_the_mocap_arm(arm):
    import torch

    from robogym_b200 import build, engine

    build.build()
    cm, blob = arm
    states, after, om, _ = rollout(blob, 60)
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, len(states), 5, outputs=("ncon", "warn", "body_xpos"))
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.qacc_warmstart.copy_(f(4))
    sim.mocap_pos.copy_(f(5)); sim.mocap_quat.copy_(f(6))
    sim.step()
    torch.cuda.synchronize()
    q, v = sim.qpos.cpu().numpy(), sim.qvel.cpu().numpy()
    eq = np.array([np.abs(q[k] - after[k][0]).max() for k in range(len(states))])
    ev = np.array([np.abs(v[k] - after[k][1]).max() for k in range(len(states))])
    assert int(sim.warn.max()) == 0
    assert eq.max() < 5e-5 and ev.max() < 1e-2, (eq.max(), ev.max())
    # the mocap body is where the data says it is
    mb = cm.name2id("body", "mocap")
    assert np.abs(sim.body_xpos[:, mb].cpu().numpy() - np.stack([s[5][0] for s in states])).max() < 1e-6


def _oracle_engine():
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "stubs"))
    from oracle_engine import OracleEngine

    return OracleEngine


def _shim_sim(factory):
    import robogym_b200.mujoco_py_shim as shim

    shim.set_engine_factory(factory)
    try:
        return shim.MjSim(shim.load_model_from_xml(MOCAP_ARM), nsubsteps=5)
    finally:
        shim.set_engine_factory(None)


def _drive_like_gym_mocap_set_action(sim, n):
    """gym.envs.robotics.utils.mocap_set_action / reset_mocap2body_xpos as the reference's MocapSolver uses them
    (robogym/robot/control/tcp/mocap_solver.py:41-46): relative moves of data.mocap_pos / mocap_quat, then sim.step()."""
    tcp_track = []
    for k in range(n):
        sim.data.set_mocap_pos("mocap", sim.data.get_mocap_pos("mocap") + [-0.001, 0.0005, -0.0005])
        sim.data.mocap_quat[0] = [1, 0, 0, 0]
        sim.data.ctrl[0] = 0.01
        sim.step()
        tcp_track.append((sim.data.get_body_xpos("tcp").copy(), sim.data.get_body_xvelp("tcp").copy(), sim.data.get_body_xvelr("tcp").copy()))
    return tcp_track


def test_shim_mocap_accessors_on_the_oracle_engine():
    """data.set_mocap_pos / mocap_quat, get_body_xvelp / get_body_xvelr (robogym/robot/ur16e/mujoco/
    joint_controlled_arm.py:31-33) through the mujoco_py look-alike, with the oracle as the engine (CPU tier)."""
    sim = _shim_sim(_oracle_engine())
    assert sim.data.mocap_pos.shape == (2, 3) and np.allclose(sim.data.get_mocap_pos("hook"), [-0.3, 0.2, 0.6])
    with pytest.raises(ValueError):
        sim.data.set_mocap_pos("tcp", [0, 0, 0])
    track = _drive_like_gym_mocap_set_action(sim, 80)
    # the tool is dragged along (a 4-joint arm cannot keep the commanded orientation while it bends, so the soft weld trades
    # position against orientation: centimetres of lag, not millimetres)
    want = sim.data.get_mocap_pos("mocap") + [0, 0, -0.05]
    assert np.abs(track[-1][0] - want).max() < 0.05 and track[-1][0][0] < 0.41 and track[-1][1][0] < -0.03
    # body velocity output = finite difference of the body position
    fd = (track[-1][0] - track[-2][0]) / (5 * 0.002)
    assert np.abs(fd - track[-1][1]).max() < 0.01
    sim.reset()
    assert np.allclose(sim.data.get_mocap_pos("mocap"), [0.45, 0.0, 0.55])


@pytest.mark.gpu
def test_shim_mocap_accessors_on_the_cuda_engine_match_the_oracle_engine():
    import torch

    from robogym_b200 import build

    assert torch.cuda.is_available()
    build.build()
    cuda = _shim_sim(None)
    ora = _shim_sim(_oracle_engine())
    assert type(cuda._rg_engine).__name__ == "CudaEngine"
    a, b = _drive_like_gym_mocap_set_action(cuda, 60), _drive_like_gym_mocap_set_action(ora, 60)
    for (xa, va, wa), (xb, vb, wb) in zip(a, b):
        assert np.abs(xa - xb).max() < 5e-4 and np.abs(va - vb).max() < 2e-2 and np.abs(wa - wb).max() < 0.1   # free-running, fp32 vs fp64
