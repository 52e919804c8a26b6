"""Generate robogym_b200/assets/rearrange_ycb8_tcp.rgm (+ names) -- build container only (needs /root/reference).

The main simulation of the UNMODIFIED reference environment `robogym.envs.rearrange.ycb.make_env` (8 YCB mesh objects drawn by
the environment itself (first starting_seed whose placement succeeds), ControlMode.TCP_ROLL_YAW + TcpSolverMode.MOCAP_IK, default cascaded-PI arm
calibration), compiled by the shim when the environment resets: BASELINE.json configs[4] as the reference builds it.  Its solver
simulation is the same arm as the blocks environment's (robogym_b200/assets/rearrange_solver_arm.rgm)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
REF = os.environ.get("ROBOGYM_REFERENCE", "/root/reference")
for p in (os.path.join(ROOT, "tests", "stubs"), REF, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import robogym_b200.mujoco_py_shim as shim

    shim.install()
    from oracle_engine import OracleEngine

    shim.set_engine_factory(OracleEngine)
    from robogym.envs.rearrange.ycb import make_env
    from robogym.robot.robot_interface import ControlMode, TcpSolverMode

    # eight objects crowd the reference's placement area: its rejection sampling gives up for some seeds; the first seed whose
    # reset goes through is used (the bench re-places the objects on a grid anyway)
    from robogym.utils.env_utils import InvalidSimulationError

    for seed in range(20):
        env = make_env(parameters=dict(n_random_initial_steps=0, simulation_params=dict(num_objects=8, max_num_objects=8),
                                       robot_control_params=dict(control_mode=ControlMode.TCP_ROLL_YAW, tcp_solver_mode=TcpSolverMode.MOCAP_IK, max_position_change=0.1)),
                       constants=dict(stabilize_objects=False), starting_seed=seed)
        try:
            env.reset()
            break
        except InvalidSimulationError:
            continue
    print("starting_seed", seed)
    u = env.unwrapped
    mj = u.mujoco_simulation.mj_sim
    solver = u.robot.robots[0].controller_arm.mj_sim
    ref_solver = open(os.path.join(ROOT, "robogym_b200", "assets", "rearrange_solver_arm.rgm"), "rb").read()
    assert solver.model._cm.blob() == ref_solver, "the YCB environment's solver arm differs from the blocks environment's"
    assets = os.path.join(ROOT, "robogym_b200", "assets")
    blob = mj.model._cm.blob()
    open(os.path.join(assets, "rearrange_ycb8_tcp.rgm"), "wb").write(blob)
    json.dump(mj.model._cm.names, open(os.path.join(assets, "rearrange_ycb8_tcp.names.json"), "w"))
    groups = [(os.path.basename(os.path.dirname(g.mesh_files[0])), g.count) if hasattr(g, "mesh_files") else str(g) for g in u.mujoco_simulation.object_groups]
    m = mj.model._m
    print("objects", groups)
    print("nq/nv/nu", m["nq"], m["nv"], m["nu"], "ngeom", m["ngeom"], "npair", m["npair"], "blob", len(blob))


if __name__ == "__main__":
    main()
