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
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
