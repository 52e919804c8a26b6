"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sharding + max-over-ranks reduction
that bench.py uses (environments shard by batch index, no data-path collective)."""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["RG_ROOT"])
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = bench.shard_range(8192 * world, rank, world)
assert hi - lo == 8192 and lo == rank * 8192
assert bench.rank_seed(1234, rank) != bench.rank_seed(1234, (rank + 1) % world)
t = bench.max_over_ranks(float(rank + 1), dist, torch.device("cpu"))
assert t == float(world), t
tot = bench.sum_over_ranks(8192.0, dist, torch.device("cpu"))
assert tot == 8192.0 * world
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
'''


def test_gloo_two_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, RG_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert "DIST_OK" in out.stdout, out.stdout + out.stderr
