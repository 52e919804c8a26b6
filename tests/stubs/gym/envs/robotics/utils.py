"""Stand-in for gym.envs.robotics.utils (gym is not installed in this image; test infrastructure): the three mocap helpers
robogym's MocapSolver calls (robogym/robot/control/tcp/mocap_solver.py:52-57), written from their documented behaviour:

* reset_mocap_welds(sim): every weld constraint's relative pose becomes the identity, then sim.forward();
* reset_mocap2body_xpos(sim): every mocap body jumps to the pose of the body it is welded to;
* mocap_set_action(sim, action): action = nmocap x (dx, dy, dz, dqw, dqx, dqy, dqz); mocap bodies are first re-seated on
  their welded bodies, then moved by the deltas (position and quaternion deltas are ADDED, as gym does)."""
import numpy as np

EQ_WELD = 1


def reset_mocap_welds(sim):
    m = sim.model
    if m.nmocap > 0 and m.neq > 0:
        for i in range(m.neq):
            if m.eq_type[i] == EQ_WELD:
                m.eq_data[i, :] = [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]
    sim.forward()


def reset_mocap2body_xpos(sim):
    m = sim.model
    for i in range(m.neq):
        if m.eq_type[i] != EQ_WELD:
            continue
        b1, b2 = int(m.eq_obj1id[i]), int(m.eq_obj2id[i])
        k, body = int(m.body_mocapid[b1]), b2
        if k == -1:
            k, body = int(m.body_mocapid[b2]), b1
        assert k != -1, "weld without a mocap body"
        sim.data.mocap_pos[k][:] = sim.data.body_xpos[body]
        sim.data.mocap_quat[k][:] = sim.data.body_xquat[body]


def mocap_set_action(sim, action):
    n = sim.model.nmocap
    if n > 0:
        a = np.asarray(action, dtype=np.float64)[:n * 7].reshape(n, 7)
        reset_mocap2body_xpos(sim)
        sim.data.mocap_pos[:] = sim.data.mocap_pos + a[:, :3]
        sim.data.mocap_quat[:] = sim.data.mocap_quat + a[:, 3:]


def ctrl_set_action(sim, action):
    raise NotImplementedError("gym.envs.robotics.utils.ctrl_set_action is not used by robogym's step path")


robot_get_obs = ctrl_set_action
