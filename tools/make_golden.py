"""Generate tests/golden/*.npz from the REFERENCE's own pure-Python pieces (build container only).

* hand_fk.npz: robogym/mujoco/forward_kinematics.py (ForwardKinematics.compute via
  robogym/robot/shadow_hand/hand_forward_kinematics.py:21-60) evaluated at seeded random joint
  angles inside the XML joint ranges: the reference-owned oracle for the kinematics stage
  (robogym/robot/shadow_hand/test/test_mujoco_hand.py:19-41 compares exactly these quantities
  against MuJoCo to 1e-6).
* locked_expectations.json: literal expectations copied from the reference's tests (joint order
  robogym/envs/dactyl/tests/test_locked.py:17-52, cube mass :59-63, ctrlrange table
  robogym/robot/shadow_hand/hand_interface.py:153-174).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import compose_reference_xml as ref  # noqa: E402

ref.mujoco_xml_cls()
from robogym.robot.shadow_hand import hand_forward_kinematics as hfk  # noqa: E402
from robogym.robot.shadow_hand import hand_interface  # noqa: E402
from robogym_b200 import mjcf  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    cm = mjcf.compile_mjcf(ref.locked_xml())
    jr = cm.m["jnt_range"].reshape(-1, 2)
    hand = [cm.name2id("joint", "robot0:" + j) for j in hand_interface.JOINTS]
    rng = np.random.RandomState(20260923)
    q = np.array([[rng.uniform(*jr[j]) for j in hand] for _ in range(64)])
    q[0] = 0.0
    rel = np.array([hfk.compute_forward_kinematics_fingertips(x) for x in q])
    np.savez(os.path.join(OUT, "hand_fk.npz"), joint_names=np.array(hand_interface.JOINTS), qpos=q, fingertips_rel=rel,
             reference_sites=np.array(hfk.REFERENCE_SITE_NAMES), fingertip_sites=np.array(hfk.FINGERTIP_SITE_NAMES))
    exp = dict(
        joint_order=["cube:cube_tx", "cube:cube_ty", "cube:cube_tz", "cube:cube_rot", "target:cube_tx", "target:cube_ty",
                     "target:cube_tz", "target:cube_rot"] + ["robot0:" + j for j in hand_interface.JOINTS],
        cube_mass=0.078, cube_mass_tol=1e-3,
        actuators=list(hand_interface.ACTUATORS),
        ctrlrange_lower=list(map(float, hand_interface.ACTUATOR_CTRLRANGE_LOWER_BOUND)),
        ctrlrange_upper=list(map(float, hand_interface.ACTUATOR_CTRLRANGE_UPPER_BOUND)),
        dims=dict(nq=38, nv=36, nu=20, nbody=31, ngeom=65, ntendon=12, npair=1243),
    )
    json.dump(exp, open(os.path.join(OUT, "locked_expectations.json"), "w"), indent=1)
    # the reference's recorded real-MuJoCo state of four stacked blocks (the only settled multi-body fixture of its kind in
    # the repo: robogym/envs/rearrange/holdouts/states/physics_tests/block_stacking4/) with the scene constants that its
    # stability test runs it under (holdouts/tests/test_stability.py:215-261, holdouts/configs/physics_tests/*.jsonnet,
    # materials/painted_wood.jsonnet + base.libsonnet, assets/xmls/primitives/box.xml, robot/ur16e/base.xml:3-22)
    st = np.load(os.path.join(ref.REF, "robogym/envs/rearrange/holdouts/states/physics_tests/block_stacking4/initial_state_4_blocks_stacked.npz"))
    stack = dict(
        obj_pos=st["obj_pos"].tolist(), obj_quat=st["obj_quat"].tolist(),
        block_half_size=0.025, density=720.0, friction=[0.85, 0.25, 0.001], condim=6, margin=0.00005, solref=[-4000.0, -200.0],
        joint_damping=0.01, joint_armature=0.001,
        table=dict(pos=[1.4508, 0.773, 0.453], half_size=[0.6075, 0.7655, 0.03324], solimp=[0.99, 0.999, 0.001], solref=[-50000.0, -100.0]),
        option=dict(timestep=0.002, substeps=20, cone="elliptic", impratio=10),
        stability=dict(env_steps=50, max_linear_speed=0.3, note="test_stability.py:215-230: block_stacking4 must stay below 0.3 m/s for 50 env-steps"),
        doc_resting_z=dict(value=0.51167315, block_half_size=0.0254, source="docs/env_param_interface.md:32-38 (default material, blocks env)"),
    )
    json.dump(stack, open(os.path.join(OUT, "block_stack4.json"), "w"), indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
