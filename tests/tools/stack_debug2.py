"""(test tooling) per-env-step block speeds of the 4-block stack on the CUDA engine."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import torch
from robogym_b200 import build, engine, mjcf
from test_blocks import stack_xml
cm = mjcf.compile_mjcf(stack_xml()); blob = cm.blob()
model = engine.DeviceModel(blob, 0)
sim = engine.BatchedSim(model, 1, 20, outputs=("ncon", "warn"), debug=True, contact_capacity=48, row_capacity=16, dofs_per_contact=12)
for k in range(14):
    sim.step(final_forward=0); torch.cuda.synchronize()
    v = sim.qvel.cpu().numpy()[0].reshape(-1, 6)
    print(k, "speeds", np.round(np.linalg.norm(v[:, :3], axis=1), 3), "ang", np.round(np.linalg.norm(v[:, 3:], axis=1), 2), "ncon", int(sim.ncon[0]), "niter", sim.dbg_view(0)["niter"], "z", np.round(sim.qpos.cpu().numpy()[0].reshape(-1, 7)[:, 2], 4))
