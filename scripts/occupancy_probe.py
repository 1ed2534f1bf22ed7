"""Kernel time of the persistent optimizer on the bench block (or its candidates with <= maxN pieces), for A/B builds
of the library (ALORE_B200_LIB=<variant .so>).  usage: python scripts/occupancy_probe.py [maxN] [n]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200.ms_planner import DeviceBatch
from test_esdf_gpu import make_sdf
import bench
maxN = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8320
ctx = alore.Context(0)
prm = alore.default_params()
geom, grid = bench.build_world()
m = make_sdf(ctx, 2048, 2048, 0.05, grid)
m.updateESDF2d()
pts = bench.way_points(m.geom(), grid, m.distance_buffer_all_)
cands = bench.candidates_from_points(pts, 0, n)
N = np.diff(cands.piece_off)
keep = np.nonzero(N <= maxN)[0]
if keep.size < cands.B:
    cands = cands.subset(keep)
print(f"lib {os.environ.get('ALORE_B200_LIB', 'default')}: {cands.B} candidates, Nmax {int(np.diff(cands.piece_off).max())}")
db = DeviceBatch(ctx, cands)
ms = []
for it in range(4):
    db.run(prm)
    r = db.download()
    ms.append(db.kernel_ms())
print("kernel ms per run (first = cold order):", [round(x, 1) for x in ms], "evals", int(r.evals.sum()), "cost sum", repr(float(np.sum(r.cost[np.isfinite(r.cost)]))))
