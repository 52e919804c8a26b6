"""(Test tooling: compares the engine with the oracle -- lives under tests/ because only tests may load oracle/.)
First-light check on a B200: teacher-forced parity of the CUDA engine vs the fp64 oracle on
states sampled along an oracle rollout, then a throughput probe at batch 8192."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import pyoracle  # noqa: E402
from robogym_b200 import build, engine  # noqa: E402

build.build()
blob = open(os.path.join(ROOT, "robogym_b200/assets/dactyl_locked.rgm"), "rb").read()
om = pyoracle.OracleModel(blob)
d = pyoracle.OracleData(om)
nu, nq, nv = om.dim("nu"), om.dim("nq"), om.dim("nv")
cr = om.field("actuator_ctrlrange").reshape(-1, 2)
rng = np.random.RandomState(1)
d.ctrl[:] = cr.mean(1)
for _ in range(20):
    d.env_step(10)
K = int(os.environ.get("K", 64))
states, after = [], []
for s in range(K):
    a = rng.uniform(-1, 1, nu)
    d.ctrl[:] = np.clip(d.ctrl + 0.3 * a * (cr[:, 1] - cr[:, 0]) / 2, cr[:, 0], cr[:, 1])
    states.append((d.qpos.copy(), d.qvel.copy(), d.ctrl.copy(), d.userdata[:3 * nu].copy(), d.qacc_warmstart.copy()))
    d.env_step(10)
    after.append((d.qpos.copy(), d.qvel.copy(), int(d.ncon[0])))

model = engine.DeviceModel(blob, 0)
sim = engine.BatchedSim(model, K, 10, outputs=("site_xpos", "act_force", "ncon", "warn"), debug=True)
print("launch", sim.launch_info(), "scratch bytes/env", model.scratch_bytes)
f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device="cuda")
sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.pid.copy_(f(3)); sim.qacc_warmstart.copy_(f(4))
sim.step()
torch.cuda.synchronize()
q = sim.qpos.cpu().numpy(); v = sim.qvel.cpu().numpy()
iq = list(range(0, 7)) + list(range(14, nq)); iv = list(range(0, 6)) + list(range(12, nv))
eq = np.array([np.abs(q[k][iq] - after[k][0][iq]).max() for k in range(K)])
ev = np.array([np.abs(v[k][iv] - after[k][1][iv]).max() for k in range(K)])
print("parity: median dq %.2e dv %.2e | 90%% dq %.2e dv %.2e | max dq %.2e dv %.2e" % (np.median(eq), np.median(ev), np.percentile(eq, 90), np.percentile(ev, 90), eq.max(), ev.max()))
print("ncon gpu", sim.ncon.cpu().numpy()[:16], "oracle", [a[2] for a in after[:16]], "warn", sim.warn.cpu().numpy().max())

# throughput probe
N = int(os.environ.get("NENV", 8192))
sim2 = engine.BatchedSim(model, N, 10)
idx = torch.arange(N, device="cuda") % K
sim2.qpos.copy_(sim.qpos[idx]); sim2.qvel.copy_(sim.qvel[idx]); sim2.ctrl.copy_(sim.ctrl[idx]); sim2.pid.copy_(sim.pid[idx])
lo = torch.tensor(cr[:, 0], dtype=torch.float32, device="cuda"); hi = torch.tensor(cr[:, 1], dtype=torch.float32, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(1234)
for it in range(3):
    sim2.ctrl.copy_(lo + (hi - lo) * torch.rand(N, nu, device="cuda", generator=g))
    sim2.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = int(os.environ.get("STEPS", 10))
e0.record()
for it in range(steps):
    sim2.ctrl.copy_(lo + (hi - lo) * torch.rand(N, nu, device="cuda", generator=g))
    sim2.step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps(dict(nenv=N, ms_per_env_step=ms, env_steps_per_s=N / ms * 1e3, launch=sim2.launch_info(), warn_max=int(sim2.warn.max().item()), ncon_mean=float(sim2.ncon.float().mean().item()))))

# stage profile (only meaningful with RG_LIB=...librobogym_b200_prof.so)
g = sim.dbg_view(0)
print("dbg env0: ncon", g["ncon"], "nel", g["nel"], "niter", g["niter"])
import os as _os
if "prof" not in _os.environ.get("RG_LIB", ""):
    raise SystemExit(0)          # the stage counters exist only in the -DRG_PROFILE build
prof = sim.dbg.cpu().numpy()[:, -16:]
names = ["kin", "massm", "bias", "tendon", "forces", "collide", "mkcon", "solve", "euler", "col:A-sphere", "col:B-obb", "col:C-narrow+write", "col:C-rounds", "-", "-", "col:loop"]
if "prof2" in _os.environ.get("RG_LIB", ""):
    names = names[:9] + ["sol:init", "sol:gradient", "sol:H-assembly", "sol:cholesky", "sol:tri-solves", "sol:linesearch", "sol:step+update"]
tot = prof[:, :9].sum(1).mean()
print("stage cycles per env-step (mean over envs, lane0 clock64):")
for i, nme in enumerate(names):
    if nme == '-': continue
    print("  %-8s %10.0f  %5.1f%%" % (nme, prof[:, i].mean(), 100 * prof[:, i].mean() / max(tot, 1)))
print("  total %.0f cycles/env-step/warp" % tot)
niters = np.array([sim.dbg_view(k)["niter"] for k in range(min(K, 32))])
print("newton iters (final forward)", niters)
