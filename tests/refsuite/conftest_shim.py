"""pytest plugin used when running the REFERENCE's own test files against the shim (tests/test_reference_suite.py):
installs robogym_b200.mujoco_py_shim as `mujoco_py` with the fp64 oracle engine (CPU tier)."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "stubs"), os.environ.get("ROBOGYM_REFERENCE", "/root/reference")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# names numpy 2 removed that the reference's 2020 TEST files still use (np.int in test_rearrange_robots.py:336, np.alltrue in
# robot/utils/tests/test_reach_helper.py:40, np.Inf in the holdout configs): restored for the test run only
for _name, _val in (("int", int), ("alltrue", np.all), ("Inf", np.inf)):
    if not hasattr(np, _name):
        setattr(np, _name, _val)

import robogym_b200.mujoco_py_shim as shim  # noqa: E402

shim.install()
if os.environ.get("RG_SHIM_ENGINE", "oracle") == "oracle":
    from oracle_engine import OracleEngine  # noqa: E402

    shim.set_engine_factory(OracleEngine)

# The UR16e's default joint calibration drives the arm with mujoco-py's cascaded-PI controller (robogym/robot/robot_interface.py:61-63):
# the reference's tests run with that default, untouched.  RG_REFSUITE_ARM_CALIBRATION=pid forces the reference's second calibration
# (plain PID) instead -- used once by tests/test_reference_suite.py to show that the impulse-response fixture tells the two apart.
if os.environ.get("RG_REFSUITE_ARM_CALIBRATION", "default") == "pid":
    import attr  # noqa: E402
    import robogym.robot.ur16e.mujoco.simulation.base as _arm_sim  # noqa: E402

    _orig_make_robot_xml = _arm_sim.ArmSimulationInterface.make_robot_xml.__func__

    def _make_robot_xml(cls, xml, robot_control_params):
        # the one place the calibration is read (robogym/robot/ur16e/mujoco/simulation/base.py:97)
        if robot_control_params.arm_joint_calibration_path != "pid":
            robot_control_params = attr.evolve(robot_control_params, arm_joint_calibration_path="pid")
        return _orig_make_robot_xml(cls, xml, robot_control_params)

    _arm_sim.ArmSimulationInterface.make_robot_xml = classmethod(_make_robot_xml)
