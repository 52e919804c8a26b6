"""Throughput of the full batched environment loop (robogym_b200.locked_env.BatchedLockedEnv): action -> ctrl,
physics, goal reward, multi-goal bookkeeping, drop handling, pooled auto-reset, observation -- everything on the
device.  Usage: python tools/env_bench.py [nenv] [steps]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from robogym_b200 import build  # noqa: E402
from robogym_b200.locked_env import make_cuda_env  # noqa: E402

build.build()
nenv = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = {}
for label, kw in (("reference_exact", {}), ("single_forward", dict(observe_forwards=0))):
    env = make_cuda_env(nenv, seed=0, **kw)
    env.reset()
    gen = torch.Generator(device=env.device)
    gen.manual_seed(1)
    acts = [torch.rand(nenv, 20, device=env.device, generator=gen) * 2 - 1 for _ in range(steps + 5)]
    for a in acts[:5]:
        env.step(a)
    torch.cuda.synchronize()
    ndone = 0
    t0 = time.perf_counter()
    for a in acts[5:]:
        obs, rew, done, info = env.step(a)
        ndone += done.sum()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[label] = dict(env_steps_per_s=nenv * steps / dt, ms_per_step=1e3 * dt / steps, forwards_per_step=env.final_forward,
                      episodes_finished=int(ndone), pool_generated=env.pool.generated, pool_rejected=env.pool.rejected,
                      on_palm=float(env.fac.on_palm(env.sim.site_xpos).float().mean()), warn=int(env.sim.warn.max()))
    del env
print(json.dumps(dict(nenv=nenv, steps=steps, **out)))
