"""Generate tests/golden/rearrange_reset.json + robogym_b200/assets/rearrange_blocks5_env.rgm (build container only).

Runs the UNMODIFIED reference environment `robogym.envs.rearrange.blocks_train.make_env` (dual-sim MOCAP_IK controller, the default
cascaded-PI arm calibration) on the mujoco_py shim with the fp64 oracle as the engine -- exactly the example of the reference's documentation
(docs/env_param_interface.md:12-38), whose printed observation holds the only real-MuJoCo numbers of this scene:
block heights 0.51167315.  `stabilize_objects` (robogym/envs/rearrange/common/utils.py:76-92: object damping 1e-3, then 100
env-steps = 2000 mj_steps + a forward after every 20) is where that number is produced, so the fixture is the simulator state
right after `set_object_damping(1e-3)`, the compiled model of the environment's main sim at that moment, and the heights the
environment reports afterwards.  tests/test_rearrange_reset_pin.py replays it on the oracle, the emulated kernel and CUDA.
"""
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


def main():
    import robogym_b200.mujoco_py_shim as shim
    from robogym_b200 import modelblob

    shim.install()
    from oracle_engine import OracleEngine

    shim.set_engine_factory(OracleEngine)
    import robogym.envs.rearrange.common.base as base
    import robogym.envs.rearrange.common.utils as U

    cap = {}
    orig = U.stabilize_objects

    def spy(sim, n_steps=100):
        d, m = sim.mj_sim.data, sim.mj_sim.model
        damping = sim.get_object_damping()
        sim.set_object_damping(1e-3)
        nu = m.nu
        cap.update(blob=m._cm.blob(), names=m._cm.names, nsub=int(sim.mj_sim.nsubsteps), nsteps=int(n_steps),
                   qpos=d.qpos.copy(), qvel=d.qvel.copy(), ctrl=d.ctrl.copy(), pid=d.userdata[:modelblob.pid_stride(m._m) * nu].copy(), warm=d.qacc_warmstart.copy(),
                   mocap_pos=d.mocap_pos.copy(), mocap_quat=d.mocap_quat.copy(),
                   obj_qposadr=[int(m.get_joint_qpos_addr(f"object{i}:joint")[0]) for i in range(sim.num_objects)])
        sim.set_object_damping(damping)          # hand the simulation back untouched, then let the reference do its thing
        orig(sim, n_steps)
        cap["qpos_after"] = d.qpos.copy()

    base.stabilize_objects = spy
    from robogym.envs.rearrange.blocks_train import make_env

    env = make_env(parameters={"simulation_params": {"num_objects": 5, "max_num_objects": 8}})
    obs = env.reset()
    z_obs = [float(obs["obj_pos"][i][2]) for i in range(5)]
    out = {k: np.asarray(cap[k]).tolist() for k in ("qpos", "qvel", "ctrl", "pid", "warm", "mocap_pos", "mocap_quat", "qpos_after")}
    out.update(nsub=cap["nsub"], nsteps=cap["nsteps"], obj_qposadr=cap["obj_qposadr"], obs_obj_z=z_obs,
               documented_z=0.51167315, documented_at="docs/env_param_interface.md:32-38",
               note="state of the environment's main sim right after set_object_damping(1e-3) inside stabilize_objects; obs_obj_z = obj_pos[:, 2] "
                    "of the observation env.reset() returns (reference env code on the shim, oracle engine)")
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "rearrange_reset.json"), "w"), indent=1)
    assets = os.path.join(ROOT, "robogym_b200", "assets")
    open(os.path.join(assets, "rearrange_blocks5_env.rgm"), "wb").write(cap["blob"])
    json.dump(cap["names"], open(os.path.join(assets, "rearrange_blocks5_env.names.json"), "w"))
    print("obs z", z_obs, "blob", len(cap["blob"]))


if __name__ == "__main__":
    main()
