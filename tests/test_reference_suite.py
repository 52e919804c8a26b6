"""Run the REFERENCE'S OWN test files, unmodified, against `robogym_b200.mujoco_py_shim` (the drop-in
boundary of SURVEY.md 8(b)).  Build-container only: needs /root/reference.  The engine behind the shim
is the fp64 oracle here (CPU tier); the same shim drives the CUDA engine on a GPU (tests/test_gpu_shim.py).

`gym`, `mock`, `pycuber`, `collision`, `_jsonnet`, ... are not installed in this image: tests/stubs holds minimal stand-ins
(test infrastructure, not part of the package)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.environ.get("ROBOGYM_REFERENCE", "/root/reference")

pytestmark = [pytest.mark.needs_reference, pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "robogym")), reason="needs /root/reference")]

CASES = [
    # (reference test file, -k expression, expected number of passing tests)
    ("robogym/envs/dactyl/tests/test_locked.py", None, 6),
    ("robogym/robot/shadow_hand/test/test_mujoco_hand.py", None, 4),
    ("robogym/envs/dactyl/tests/test_reach.py", None, 1),
    ("robogym/robot/shadow_hand/test/test_hand_interface.py", None, 7),
    ("robogym/envs/dactyl/tests/test_cube_utils.py", None, 3),
    ("robogym/wrappers/tests/test_dactyl.py", None, 1),
    ("robogym/tests/test_robot_env.py", None, 1),
    # test_remove_elem compares XML attribute order of the reference's own pure-Python composer under py3.12
    ("robogym/mujoco/test/test_mujoco_utils.py", "not remove_elem", 5),
    ("robogym/envs/tests/test_wrapper_compositions.py", None, 1),
    ("robogym/randomization/tests/test_randomization.py", None, 4),
    # the other tests of this file reset the full Rubik's cube env, which needs the real pycuber package
    ("robogym/wrappers/tests/test_randomizations.py", "randomize_obs_wrapper or replace_cube_obs_vision_wrapper", 2),
    # ---- rearrange (BASELINE configs[3]): the reference's environments on the shim, dual-sim MOCAP_IK controller included.
    # Run with the reference's "pid" arm calibration (tests/refsuite/conftest_shim.py): the default "cascaded_pi" controller's
    # law is not in the reference tree and the engines refuse such models.  Not selected:
    # rendering (hide_geoms), and test_mocap_ik_impulse_response, whose expected displacements belong to the cascaded-PI arm.
    ("robogym/envs/rearrange/tests/test_robot_polymorphism.py", None, 7),
    ("robogym/envs/rearrange/tests/test_placement.py", "not ycb", 4),
    ("robogym/envs/rearrange/tests/test_rearrange_sim.py", "not impulse and not hide_geoms", 6),
    ("robogym/envs/rearrange/tests/test_multi_goals_env.py", None, 6),
    ("robogym/envs/rearrange/tests/test_object_creation.py", None, 3),
    ("robogym/envs/rearrange/tests/test_goal_generation.py", None, 6),
    # Also green on the shim but too slow for this tier with the dense fp64 oracle as the engine (run by hand, same command):
    # test_placement.py -k ycb (8 tests, 32 YCB objects = 200 dofs: 20 min), test_object_rotation.py (12 tests, 6 min),
    # test_rearrange_envs.py (20 of 30: the rest need the holdout configs' full Jsonnet, numpy < 2 (`np.Inf`), or the cascaded-PI arm).
]


@pytest.mark.parametrize("path,kexpr,npass", CASES, ids=[c[0].split("/")[-1] for c in CASES])
def test_reference_test_file_passes_on_the_shim(path, kexpr, npass, tmp_path):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "refsuite"), RG_SHIM_ENGINE="oracle")
    cmd = [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-q", "-p", "conftest_shim", f"--rootdir={tmp_path}",
           os.path.join(REF, path)]
    if kexpr:
        cmd += ["-k", kexpr]
    out = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=1200)
    tail = out.stdout[-2000:] + out.stderr[-1000:]
    assert out.returncode == 0, tail
    assert f"{npass} passed" in out.stdout, tail
