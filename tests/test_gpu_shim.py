"""The mujoco_py shim on its product engine (CUDA, batch of one) -- run on the B200 box with -m gpu.
The reference tree is not on that box, so the model comes from the committed blob."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_shim_mjsim_on_cuda_engine_matches_oracle_engine(locked_blob, locked_names):
    import torch

    assert torch.cuda.is_available()
    import robogym_b200.mujoco_py_shim as shim
    from robogym_b200 import build, mjcf

    build.build()
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "stubs"))
    from oracle_engine import OracleEngine

    sims = []
    for factory in (None, OracleEngine):
        shim.set_engine_factory(factory)
        model = shim.PyMjModel(mjcf.CompiledModel.from_blob(locked_blob, locked_names))
        sim = shim.MjSim(model, nsubsteps=10)
        shim.cymj.set_pid_control(sim.model, sim.data)
        sims.append(sim)
    shim.set_engine_factory(None)
    cuda, ora = sims
    assert type(cuda._rg_engine).__name__ == "CudaEngine"
    cr = cuda.model.actuator_ctrlrange
    rng = np.random.RandomState(0)
    hand_q = [cuda.model.get_joint_qpos_addr(n) for n in cuda.model.joint_names if n.startswith("robot0:")]
    for k in range(8):
        c = cr[:, 0] + (cr[:, 1] - cr[:, 0]) * rng.uniform(0.3, 0.7, len(cr))
        for s in (cuda, ora):
            s.data.ctrl[:] = c
            s.step()
            s.forward()
        # teacher forcing: keep the two engines on the same trajectory, compare each env-step
        dq = np.abs(cuda.data.qpos[hand_q] - ora.data.qpos[hand_q]).max()
        assert dq < 2e-3, (k, dq)
        assert abs(cuda.data.time - ora.data.time) < 1e-6
        cuda.data.qpos[:] = ora.data.qpos; cuda.data.qvel[:] = ora.data.qvel
        cuda.data.userdata[:] = ora.data.userdata; cuda.data.qacc_warmstart[:] = ora.data.qacc_warmstart
    tip = cuda.data.get_site_xpos("robot0:S_fftip")
    assert np.abs(tip - ora.data.get_site_xpos("robot0:S_fftip")).max() < 1e-3
    # in-place model edits reach the device (randomisers do this)
    cuda.model.opt.gravity[:] = [0, 0, -1.0]
    z0 = cuda.data.qpos[9]
    v0 = cuda.data.qvel[8]
    cuda.step()
    assert abs((cuda.data.qvel[8] - v0) / (10 * 0.008) + 1.0) < 1e-3 and cuda.data.qpos[9] < z0
