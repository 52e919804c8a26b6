"""Force / torque sensors (mjSENS_FORCE / mjSENS_TORQUE): the UR16e's tool flange carries one of each on site `robot0:grip`
(robogym/assets/xmls/robot/ur16e/base.xml:48-49); robogym reads them as the arm's tcp_force / tcp_torque observation and for its
force-based safety stop (robogym/robot/ur16e/mujoco/joint_controlled_arm.py:35-45,84-85).  They report the wrench between the
site's body and its parent (mj_rnePostConstraint's cfrc_int) in the site's frame: closed forms on the oracle, the kernel logic in
CPU emulation against the oracle, and (gpu) the CUDA engine against it."""
import numpy as np
import pytest

import pyemu
from helpers import oracle_pair
from robogym_b200 import mjcf, modelblob

# a two-joint arm held by position servos; the tool body carries a tray (the sensor site sits at its mount, turned 90 degrees about
# x so that site axes differ from world axes) and a free ball that rests on the tray (rolling friction keeps it there when the arm sags or swings gently)
FT_ARM = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" iterations="50" tolerance="1e-12"/>
  <size nuserdata="0" njmax="100" nconmax="20"/>
  <worldbody>
    <body name="upper" pos="0 0 0.5">
      <joint name="shoulder" type="hinge" axis="0 1 0" damping="40" armature="0.05"/>
      <geom name="g_upper" type="capsule" fromto="0 0 0 0.3 0 0" size="0.02" density="900" contype="0" conaffinity="0"/>
      <body name="fore" pos="0.3 0 0">
        <joint name="elbow" type="hinge" axis="0 0 1" damping="2" armature="0.05"/>
        <geom name="g_fore" type="capsule" fromto="0 0 0 0.2 0 0" size="0.015" density="900" contype="0" conaffinity="0"/>
        <body name="tool" pos="0.2 0 0">
          <site name="grip" pos="0 0 0" quat="0.70710678 0.70710678 0 0"/>
          <geom name="tray" type="box" pos="0.06 0 0" size="0.06 0.06 0.005" density="800" condim="6" friction="1 0.005 0.002"/>
          <body name="lip" pos="0.12 0 0.02">
            <geom name="g_lip" type="box" size="0.005 0.06 0.015" density="800" contype="0" conaffinity="0"/>
          </body>
        </body>
      </body>
    </body>
    <body name="ball" pos="0.56 0 0.5249">
      <joint name="ball_free" type="free"/>
      <geom name="ball" type="sphere" size="0.02" density="2000" condim="6" friction="1 0.005 0.002"/>
    </body>
  </worldbody>
  <actuator>
    <position name="a_shoulder" joint="shoulder" kp="20000"/>
    <position name="a_elbow" joint="elbow" kp="400"/>
  </actuator>
  <sensor>
    <force name="f_grip" site="grip"/>
    <torque name="t_grip" site="grip"/>
    <jointpos name="p_elbow" joint="elbow"/>
  </sensor>
</mujoco>
"""


@pytest.fixture(scope="module")
def arm():
    cm = mjcf.compile_mjcf(FT_ARM)
    return cm, cm.blob()


def _site_rot():
    # site quat (0.7071, 0.7071, 0, 0): 90 degrees about x: site y = world z, site z = -world y (arm at rest along world x)
    return np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0.0]])


def test_compiler_places_the_sensors(arm):
    cm, _ = arm
    assert list(cm.m["sensor_type"]) == [4, 5, 8] and list(cm.m["sensor_dim"]) == [3, 3, 1] and cm.m["nsensordata"] == 7


def test_oracle_static_readings_are_the_carried_weight_and_its_moment(arm):
    cm, blob = arm
    m = cm.m
    om, d = oracle_pair(blob)
    for _ in range(4000):
        d.step()
    d.forward()
    assert np.abs(d.qvel).max() < 5e-3 and int(d.ncon[0]) == 1      # quasi-static: the ball creeps on the soft friction rows
    tool, lip, ball = (cm.name2id("body", n) for n in ("tool", "lip", "ball"))
    g = 9.81
    # at rest the flange carries tool + lip + the ball resting on the tray: force = total weight, upwards (what the parent applies)
    masses = {b: m["body_mass"][b] for b in (tool, lip, ball)}
    com = {b: d.xipos[3 * b:3 * b + 3].copy() for b in (tool, lip, ball)}
    site = d.site_xpos[:3].copy()
    R = d.site_xmat[:9].reshape(3, 3)
    f_world = np.array([0, 0, sum(masses.values()) * g])
    t_world = sum(np.cross(com[b] - site, [0, 0, masses[b] * g]) for b in (tool, lip, ball))
    assert np.abs(d.sensordata[:3] - R.T @ f_world).max() < 2e-3, (d.sensordata[:3], R.T @ f_world)   # of 2.07 N (creep friction)
    # the ball's weight reaches the tray through a soft contact a little off its centre line: compare the moment with the contact point
    con = d.contact[:24]
    t_world_c = sum(np.cross(com[b] - site, [0, 0, masses[b] * g]) for b in (tool, lip)) + np.cross(con[1:4] - site, [0, 0, masses[ball] * g])
    assert np.abs(d.sensordata[3:6] - R.T @ t_world_c).max() < 2e-4, (d.sensordata[3:6], R.T @ t_world_c)
    assert np.abs(d.sensordata[3:6] - R.T @ t_world).max() < 1e-3
    # the arm sags a little under gravity, so compare the frame with the nominal one loosely
    assert np.abs(R - _site_rot()).max() < 0.05
    assert d.sensordata[6] == d.qpos[1]


def test_oracle_reading_includes_inertial_and_applied_forces(arm):
    cm, blob = arm
    m = cm.m
    om, d = oracle_pair(blob)
    tool, lip, ball = (cm.name2id("body", n) for n in ("tool", "lip", "ball"))
    # free fall of the whole arm (servos off, no damping would be cleaner; instead: compare with Newton on the subtree directly).
    # Push the tool with xfrc_applied and swing the arm: cfrc_int must equal sum over {tool, lip} of m (a_com + g) - applied force
    d.qpos[:2] = [0.3, -0.5]; d.qvel[:2] = [1.5, -2.0]
    d.qpos[2:5] = [5, 5, 5]                    # ball out of the way
    d.xfrc_applied[6 * tool:6 * tool + 6] = [1.0, -2.0, 0.5, 0.02, 0.01, -0.03]
    d.ctrl[:] = [0.2, 0.1]
    d.forward()
    # linear acceleration of each centre of mass by finite differences of the oracle's own kinematics
    h = 1e-6
    q0, v0, a0 = d.qpos.copy(), d.qvel.copy(), d.qacc.copy()
    pos = []
    for k in (-1, 0, 1):
        _, d2 = oracle_pair(blob)
        d2.qpos[:] = q0; d2.qpos[:2] = q0[:2] + k * h * v0[:2] + 0.5 * (k * h) ** 2 * a0[:2]
        d2.forward()
        pos.append({b: d2.xipos[3 * b:3 * b + 3].copy() for b in (tool, lip)})
    f = np.zeros(3)
    for b in (tool, lip):
        acc = (pos[2][b] - 2 * pos[1][b] + pos[0][b]) / h ** 2
        f += m["body_mass"][b] * (acc + [0, 0, 9.81])
    f -= d.xfrc_applied[6 * tool:6 * tool + 3]
    R = d.site_xmat[:9].reshape(3, 3)
    assert np.abs(d.sensordata[:3] - R.T @ f).max() < 2e-3, (d.sensordata[:3], R.T @ f)


def _rollout(blob, n, nsub=5):
    om, d = oracle_pair(blob)
    rng = np.random.RandomState(2)
    states, after = [], []
    for k in range(n):
        d.ctrl[:] = [0.0, 0.15 * np.sin(0.13 * k) + rng.uniform(-0.002, 0.002)]   # the tray swings sideways, level
        if k == 30:
            d.xfrc_applied[6 * 3:6 * 3 + 3] = [0.5, 0.3, -0.4]      # the tool is pushed from then on
        states.append((d.qpos.copy(), d.qvel.copy(), d.ctrl.copy(), np.zeros(3 * 2), d.qacc_warmstart.copy(), d.xfrc_applied.copy()))
        for _ in range(nsub):
            d.step()
        d.forward()
        after.append((d.qpos.copy(), d.sensordata[:7].copy(), int(d.ncon[0])))
    return states, after


def test_emulated_kernel_matches_oracle_readings(arm):
    cm, blob = arm
    states, after = _rollout(blob, 60)
    e = pyemu.EmuBatch(blob, cm.m, len(states))
    e.xfrc = np.zeros((len(states), cm.m["nbody"], 6), np.float32)
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st[:5]
        e.xfrc[k] = st[5].reshape(-1, 6)
    e.step(5, 1)
    want = np.stack([a[1] for a in after])
    got = e.sensordata[:, :7]
    scale = np.abs(want[:, :3]).max()
    assert scale > 2.0                                   # newtons: the tray, the lip and the ball
    ncon = np.array([a[2] for a in after])
    assert (ncon > 0).sum() > 20                          # the ball rides on the tray for much of the rollout
    err_f = np.abs(got[:, :3] - want[:, :3]).max(axis=1)
    err_t = np.abs(got[:, 3:6] - want[:, 3:6]).max(axis=1)
    assert np.median(err_f) < 2e-3 and np.median(err_t) < 2e-4, (np.median(err_f), np.median(err_t))
    assert (err_f < 0.05 * scale).mean() > 0.95           # a contact that switches within fp32 noise moves the ball's share
    assert np.abs(got[:, 6] - want[:, 6]).max() < 1e-5


@pytest.mark.gpu
def test_cuda_matches_oracle_readings(arm):
    import torch

    from robogym_b200 import build, engine

    build.build()
    cm, blob = arm
    states, after = _rollout(blob, 60)
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, len(states), 5, outputs=("sensordata", "warn"))
    sim.enable_xfrc()
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.qacc_warmstart.copy_(f(4))
    sim.xfrc_applied.copy_(f(5).reshape(sim.xfrc_applied.shape))
    sim.step()
    torch.cuda.synchronize()
    want = np.stack([a[1] for a in after])
    got = sim.sensordata.cpu().numpy()[:, :7]
    err_f = np.abs(got[:, :3] - want[:, :3]).max(axis=1)
    err_t = np.abs(got[:, 3:6] - want[:, 3:6]).max(axis=1)
    assert int(sim.warn.max()) == 0
    assert np.median(err_f) < 5e-3 and np.median(err_t) < 5e-4, (np.median(err_f), np.median(err_t))
    assert (err_f < 0.05 * np.abs(want[:, :3]).max()).mean() > 0.95


# the sensor body below a FREE base and a BALL joint: the axes of a ball / free joint's rotational dofs turn with the velocity that
# precedes them (mj_comVel), which the one-lane subtree walk of the kernel has to reproduce from the dof bit masks
FT_TUMBLER = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" iterations="50" tolerance="1e-12"/>
  <size nuserdata="0" njmax="100" nconmax="20"/>
  <worldbody>
    <body name="base" pos="0 0 1.0">
      <joint name="base_free" type="free"/>
      <geom name="g_base" type="box" size="0.08 0.05 0.03" density="900" contype="0" conaffinity="0"/>
      <body name="link" pos="0.1 0 0">
        <joint name="shoulder_ball" type="ball" damping="0.02"/>
        <geom name="g_link" type="capsule" fromto="0 0 0 0.2 0 0" size="0.015" density="900" contype="0" conaffinity="0"/>
        <body name="tool" pos="0.2 0 0">
          <joint name="wrist" type="hinge" axis="0 1 0" damping="0.01"/>
          <site name="grip" pos="0.01 0 0.02" quat="0.9238795 0 0.3826834 0"/>
          <geom name="g_tool" type="box" pos="0.05 0 0" size="0.05 0.03 0.01" density="1200" contype="0" conaffinity="0"/>
        </body>
      </body>
    </body>
  </worldbody>
  <sensor>
    <force name="f_grip" site="grip"/>
    <torque name="t_grip" site="grip"/>
    <force name="f_link" site="grip"/>
  </sensor>
</mujoco>
"""


def test_emulated_kernel_matches_oracle_below_free_and_ball_joints():
    cm = mjcf.compile_mjcf(FT_TUMBLER)
    blob = cm.blob()
    om, d = oracle_pair(blob)
    rng = np.random.RandomState(4)
    d.qvel[:] = rng.uniform(-2, 2, cm.m["nv"])                 # tumbling, swinging, spinning
    q = rng.normal(size=4); d.qpos[7:11] = q / np.linalg.norm(q)
    tool = cm.name2id("body", "tool")
    states, want = [], []
    for k in range(40):
        d.xfrc_applied[6 * tool:6 * tool + 6] = rng.uniform(-0.5, 0.5, 6) if k % 3 == 0 else 0.0
        states.append((d.qpos.copy(), d.qvel.copy(), d.qacc_warmstart.copy(), d.xfrc_applied.copy()))
        for _ in range(4):
            d.step()
        d.forward()
        want.append(d.sensordata[:9].copy())
    e = pyemu.EmuBatch(blob, cm.m, len(states))
    e.xfrc = np.zeros((len(states), cm.m["nbody"], 6), np.float32)
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.warm[k] = st[:3]
        e.xfrc[k] = st[3].reshape(-1, 6)
    e.step(4, 1)
    want = np.stack(want)
    assert np.abs(want[:, :3]).max() > 1.0 and np.abs(want[:, 3:6]).max() > 0.05            # centrifugal + gravity loads, not noise
    assert np.abs(e.sensordata[:, :9] - want).max() < 2e-3 * max(1.0, np.abs(want).max()), np.abs(e.sensordata[:, :9] - want).max()
    # sanity of the oracle side: in free fall (no applied force, nothing spinning) the flange transmits nothing
    _, d2 = oracle_pair(blob)
    d2.forward()
    assert np.abs(d2.sensordata[:6]).max() < 1e-9
