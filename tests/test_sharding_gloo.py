"""world_size-2 gloo test (CPU) of the multi-GPU host logic: block sharding + the (best cost, index) all-gather."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np

from alore_legged_manipulator_b200 import sharding

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from alore_legged_manipulator_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total = 1001
rng = np.random.default_rng(123)                      # same global arrays on every rank
cost = rng.uniform(10, 1000, size=total)
ok = (rng.random(total) > 0.3).astype(np.int32)
lo, hi = sharding.shard_range(rank, world, total=total)
c, i = sharding.local_best(cost[lo:hi], ok[lo:hi], lo)
gc, gi, pairs = sharding.gather_best(c, i)
valid = np.flatnonzero(ok == 1)
exp = int(valid[np.argmin(cost[valid])])
assert gi == exp and gc == cost[exp], (gi, exp)
assert len(pairs) == world and pairs[rank] == (c, i)
# a rank with no successful candidate reports (inf, -1) and never wins
c2, i2 = sharding.local_best(cost[lo:hi], np.zeros(hi - lo, np.int32) if rank == 0 else ok[lo:hi], lo)
gc2, gi2, _ = sharding.gather_best(c2, i2)
lo1, hi1 = sharding.shard_range(1, world, total=total)
v1 = np.flatnonzero(ok[lo1:hi1] == 1)
if world == 2:
    assert gi2 == lo1 + int(v1[np.argmin(cost[lo1:hi1][v1])])
dist.barrier()
if rank == 0:
    print("OK", gi, flush=True)
dist.destroy_process_group()
'''


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_partition_the_batch():
    for world in (1, 2, 3, 4, 8):
        for total in (1, 7, 16640, 65536):
            edges = [sharding.shard_range(r, world, total=total) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_range(3, 8, per_rank=2080) == (6240, 8320)
    offs = sharding.balanced_blocks(np.r_[np.full(100, 80), np.full(300, 10)], 2)
    assert offs[0] == 0 and offs[-1] == 400 and 60 < offs[1] < 80      # heavy candidates -> smaller block


def test_gather_best_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", env["MASTER_PORT"], str(script), str(ROOT)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout


def test_gather_best_single_process():
    c, i, pairs = sharding.gather_best(3.5, 42)
    assert (c, i) == (3.5, 42) and pairs == [(3.5, 42)]
