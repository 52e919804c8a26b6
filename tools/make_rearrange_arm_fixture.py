"""Generate robogym_b200/assets/rearrange_solver_arm.rgm, rearrange_blocks5_tcp.rgm and tests/golden/rearrange_arm.json (build
container only: needs /root/reference).

Runs the UNMODIFIED reference environment `robogym.envs.rearrange.blocks.make_env` with ControlMode.TCP_ROLL_YAW +
TcpSolverMode.MOCAP_IK (tool translation + roll / yaw + gripper; default cascaded-PI arm calibration) on the
mujoco_py shim with the fp64 oracle as engine, and records: the compiled models of its two simulations (main scene, solver arm),
their states right after env.reset(), a sequence of float32 actions, and the state of both simulations after every env.step.
tests/test_rearrange_arm.py replays it through robogym_b200.rearrange_arm.BatchedTcpArmController."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
REF = os.environ.get("ROBOGYM_REFERENCE", "/root/reference")
for p in (os.path.join(ROOT, "tests", "stubs"), REF, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

MAX_POSITION_CHANGE = 0.1
NSTEPS = 16


def _state(mj, w):
    d = mj.data
    return dict(qpos=d.qpos.tolist(), qvel=d.qvel.tolist(), ctrl=d.ctrl.tolist(), pid=d.userdata[:w].tolist(), warm=d.qacc_warmstart.tolist(),
                mocap_pos=np.asarray(d.mocap_pos).tolist(), mocap_quat=np.asarray(d.mocap_quat).tolist(),
                body_xpos=np.asarray(d.body_xpos).tolist(), body_xquat=np.asarray(d.body_xquat).tolist())


def main():
    import robogym_b200.mujoco_py_shim as shim
    from robogym_b200 import modelblob

    shim.install()
    from oracle_engine import OracleEngine

    shim.set_engine_factory(OracleEngine)
    from robogym.envs.rearrange.blocks import make_env
    from robogym.robot.robot_interface import ControlMode, TcpSolverMode

    out = {}
    for rce in (True, False):
        env = make_env(parameters=dict(n_random_initial_steps=0, simulation_params=dict(num_objects=5),
                                       robot_control_params=dict(control_mode=ControlMode.TCP_ROLL_YAW, tcp_solver_mode=TcpSolverMode.MOCAP_IK,
                                                                 arm_reset_controller_error=rce, max_position_change=MAX_POSITION_CHANGE)),
                       starting_seed=0)
        env.reset()
        env = env.unwrapped
        main_mj = env.mujoco_simulation.mj_sim
        arm = env.robot.robots[0]
        solver_mj = arm.controller_arm.mj_sim
        wm = modelblob.pid_stride(main_mj.model._m) * main_mj.model.nu
        ws = modelblob.pid_stride(solver_mj.model._m) * solver_mj.model.nu
        blobs = (main_mj.model._cm.blob(), solver_mj.model._cm.blob())
        names = (main_mj.model._cm.names, solver_mj.model._cm.names)
        rec = dict(reset_controller_error=rce, max_position_change=float(arm.controller_arm.max_position_change),
                   nsub_main=int(main_mj.nsubsteps), nsub_solver=int(solver_mj.nsubsteps),
                   main0=_state(main_mj, wm), solver0=_state(solver_mj, ws), actions=[], main_qpos=[], main_ctrl=[], solver_qpos=[], solver_mocap_pos=[])
        rng = np.random.RandomState(0)
        for k in range(NSTEPS):
            a = rng.uniform(-1, 1, 6).astype(np.float32)
            if k == 5:
                a[4] = 1.0                       # drive the wrist towards its range (constrain_quat_ctrl)
            env.step(a)
            rec["actions"].append([float(x) for x in a])
            rec["main_qpos"].append(main_mj.data.qpos.tolist()); rec["main_ctrl"].append(main_mj.data.ctrl.tolist())
            rec["solver_qpos"].append(solver_mj.data.qpos.tolist()); rec["solver_mocap_pos"].append(np.asarray(solver_mj.data.mocap_pos).tolist())
        out["reset_error_%s" % str(rce).lower()] = rec
        if rce:
            assets = os.path.join(ROOT, "robogym_b200", "assets")
            for stem, blob, nm in (("rearrange_blocks5_tcp", blobs[0], names[0]), ("rearrange_solver_arm", blobs[1], names[1])):
                open(os.path.join(assets, stem + ".rgm"), "wb").write(blob)
                json.dump(nm, open(os.path.join(assets, stem + ".names.json"), "w"))
            print("blobs", len(blobs[0]), len(blobs[1]))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "rearrange_arm.json"), "w"))
    print("wrote fixture;", {k: len(v["actions"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
