"""Host-side overhead of minco_plan_batch on the bench block (8320 candidates, 70 MB of results): a fresh pageable
result per call vs ONE page-locked reusable store and page-locked candidates.
python scripts/e2e_host_overhead.py [reps]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from alore_legged_manipulator_b200 import capi
from alore_legged_manipulator_b200.ms_planner import MSPlanner
from alore_legged_manipulator_b200.sdf_map import SDFmap

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = capi.Context(0)
prm = capi.default_params()
geom, grid = bench.build_world()
from alore_legged_manipulator_b200 import workloads
sys.path.insert(0, str(ROOT / "tests"))
from test_esdf_gpu import make_sdf
m = make_sdf(ctx, 2048, 2048, 0.05, grid)
m.updateESDF2d()
pts = bench.way_points(m.geom(), grid, m.distance_buffer_all_)
cands = bench.candidates_from_points(pts, 0, 8320)
pl = MSPlanner(ctx, prm, m)
pl.minco_plan_batch(cands)                       # warm (schedule prediction, allocations)
store = capi.ResultBatch(capacity=(cands.B, cands.total_pieces)).pin(ctx)
for mode in ("fresh", "store", "fresh", "store+pinned-in", "fresh"):
    if mode == "store+pinned-in":
        cands.pin(ctx)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = pl.minco_plan_batch(cands) if mode == "fresh" else pl.minco_plan_batch(cands, out=store)
        ts.append(time.perf_counter() - t0)
    print(f"{mode:16s} wall ms: " + " ".join(f"{1e3 * t:8.1f}" for t in ts))
