"""One launch of the persistent optimizer on copies of one heavy candidate (for ncu): python scripts/contention_ncu.py <warps_per_sm>"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200.ms_planner import DeviceBatch
from test_esdf_gpu import make_sdf
import bench
wps = int(sys.argv[1])
os.environ["ALORE_OPT_WARPS_PER_SM"] = str(wps)
ctx = alore.Context(0)
geom, grid = bench.build_world()
m = make_sdf(ctx, 2048, 2048, 0.05, grid)
m.updateESDF2d()
pts = bench.way_points(m.geom(), grid, m.distance_buffer_all_)
cands = bench.candidates_from_points(pts, 0, 8320)
prm = alore.default_params()
import numpy as np
N = np.diff(cands.piece_off)
if len(sys.argv) > 2:
    pick = int(sys.argv[2])
else:                                   # a long trajectory with a short optimisation: cheap to replay under ncu
    full = DeviceBatch(ctx, cands)
    full.run(prm); r = full.download(); full.close()
    ok = np.nonzero((N >= 55) & (r.evals >= 80) & (r.evals <= 140))[0]
    pick = int(ok[0])
    print("picked candidate", pick, "pieces", int(N[pick]), "evals", int(r.evals[pick]))
db = DeviceBatch(ctx, cands.subset([pick] * (wps * 148)))
db.run(prm); db.download()
print("kernel ms", db.kernel_ms())
