"""data.sensordata (SURVEY 8a' S14 / 8b): joint-position and touch sensors.  The hand has five fingertip touch pads
(robogym/assets/xmls/robot/shadowhand/assets.xml:135-142, capsule sites chain.xml:65,242); the UR16e scene adds joint
position sensors (robogym/assets/xmls/robot/ur16e/base.xml:41-46).  Force / torque sensors: tests/test_force_torque_sensors.py."""
import numpy as np
import pytest

import pyemu
from helpers import contract_states, oracle_pair
from robogym_b200 import mjcf, modelblob

PAD = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" iterations="50" tolerance="1e-12"/>
  <size nuserdata="0" njmax="100" nconmax="20"/>
  <worldbody>
    <body name="floor" pos="0 0 0"><geom name="floor" type="plane" size="2 2 1" condim="3"/></body>
    <body name="ball" pos="0 0 0.0499">
      <joint name="ball_free" type="free"/>
      <geom name="ball" type="sphere" size="0.05" density="700" condim="3"/>
      <site name="pad_low" type="capsule" size="0.02 0.01" pos="0 0 -0.045"/>
      <site name="pad_high" type="sphere" size="0.02" pos="0 0 0.045"/>
    </body>
    <body name="slider" pos="0.5 0 0.3">
      <joint name="lift" type="slide" axis="0 0 1" damping="1"/>
      <geom name="g_slider" type="box" size="0.02 0.02 0.02" density="500" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
  <sensor>
    <touch name="t_low" site="pad_low"/>
    <touch name="t_high" site="pad_high"/>
    <jointpos name="p_lift" joint="lift"/>
  </sensor>
</mujoco>
"""


def test_touch_sensor_reads_the_weight_and_jointpos_reads_qpos():
    """A ball at rest on the floor: the pad around the contact point reads m g, the pad on top reads 0 (oracle to 1e-6 N, the
    kernel logic in CPU emulation to 1e-4 N); the joint-position sensor is the joint's qpos."""
    cm = mjcf.compile_mjcf(PAD)
    blob = cm.blob()
    assert cm.m["nsensordata"] == 3 and list(cm.m["sensor_type"]) == [0, 0, 8]
    assert list(cm.m["site_type"]) == [3, 2] and np.allclose(cm.m["site_size"].reshape(-1, 3)[0, :2], [0.02, 0.01])
    mg = 700 * 4 / 3 * np.pi * 0.05 ** 3 * 9.81
    om, d = oracle_pair(blob)
    for _ in range(3000):
        d.step()
    d.forward()
    lift = cm.m["jnt_qposadr"][cm.name2id("joint", "lift")]
    assert abs(d.sensordata[0] - mg) < 1e-6 and d.sensordata[1] == 0 and d.sensordata[2] == d.qpos[lift] and d.qpos[lift] < -0.01
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, 1)
    e.qpos[0] = cm.m["qpos0"]
    e.step(3000, 1)
    assert abs(float(e.sensordata[0, 0]) - mg) < 1e-4 and float(e.sensordata[0, 1]) == 0.0
    assert float(e.sensordata[0, 2]) == float(e.qpos[0, lift])
    # turned upside down (the top pad now faces the floor) the readings swap
    q = cm.m["qpos0"].copy()
    q[3:7] = [0, 1, 0, 0]
    d.reset(); d.qpos[:] = q
    for _ in range(3000):
        d.step()
    d.forward()
    assert d.sensordata[0] == 0 and abs(d.sensordata[1] - mg) < 1e-6


def _hand_touch(blob, seeds, per_seed):
    states, after, om = contract_states(blob, seeds, per_seed)
    _, d = oracle_pair(blob)
    want = []
    for st in states:
        d.qpos[:], d.qvel[:], d.ctrl[:] = st[0], st[1], st[2]
        d.userdata[:len(st[3])] = st[3]
        d.qacc_warmstart[:] = st[4]
        d.env_step(10)
        want.append(d.sensordata[:5].copy())
    return states, np.array(want)


def test_emulated_fingertip_touch_matches_oracle_on_the_hand(locked_blob):
    """The five fingertip pads of dactyl/locked on states of the contract workload: the kernel logic reproduces which pads
    are loaded and by how much."""
    m = modelblob.unpack(locked_blob)
    assert m["nsensor"] == 5 and set(m["sensor_type"]) == {0}
    states, want = _hand_touch(locked_blob, range(400, 402), 12)
    e = pyemu.EmuBatch(locked_blob, {k: m[k] for k in modelblob.DIMS}, len(states))
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st
    e.step(10, 1)
    got = e.sensordata[:, :5]
    assert (want > 0.01).sum() >= 5, "the fixture never loads a pad"
    assert np.mean((want > 0.01) == (got > 0.01)) > 0.95
    both = (want > 0.01) & (got > 0.01)
    assert np.median(np.abs(got[both] - want[both]) / want[both]) < 0.05


@pytest.mark.gpu
def test_cuda_fingertip_touch_matches_oracle_on_the_hand(locked_blob):
    import torch

    from robogym_b200 import build, engine

    build.build()
    states, want = _hand_touch(locked_blob, range(400, 404), 16)
    model = engine.DeviceModel(locked_blob, 0)
    sim = engine.BatchedSim(model, len(states), 10, outputs=("sensordata", "warn"))
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))
    sim.step()
    torch.cuda.synchronize()
    got = sim.sensordata.cpu().numpy()[:, :5]
    assert (want > 0.01).sum() >= 10
    assert np.mean((want > 0.01) == (got > 0.01)) > 0.95
    both = (want > 0.01) & (got > 0.01)
    assert np.median(np.abs(got[both] - want[both]) / want[both]) < 0.05


def test_shim_exposes_sensordata_with_the_mujoco_py_addressing():
    """sim.data.sensordata[model.sensor_adr[id] : + model.sensor_dim[id]] with id = model.sensor_name2id(name), the access
    pattern of robogym/robot/ur16e/mujoco/joint_controlled_arm.py:35-45, on the oracle engine (CPU tier)."""
    import os
    import sys

    import robogym_b200.mujoco_py_shim as shim

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "stubs"))
    from oracle_engine import OracleEngine

    shim.set_engine_factory(OracleEngine)
    try:
        sim = shim.MjSim(shim.load_model_from_xml(PAD), nsubsteps=10)
    finally:
        shim.set_engine_factory(None)
    for _ in range(300):
        sim.step()
    sim.forward()
    sid = sim.model.sensor_name2id("t_low")
    adr, dim = sim.model.sensor_adr[sid], sim.model.sensor_dim[sid]
    mg = 700 * 4 / 3 * np.pi * 0.05 ** 3 * 9.81
    assert dim == 1 and abs(sim.data.sensordata[adr:adr + dim][0] - mg) < 1e-5
    pid = sim.model.sensor_name2id("p_lift")
    assert sim.data.sensordata[sim.model.sensor_adr[pid]] == sim.data.get_joint_qpos("lift")
