"""A/B of the two optimizer paths on the bench block: persistent warp-per-candidate
(default) vs wavefront (ALORE_OPT_WAVE=1).  usage: python scripts/opt_ab.py [n]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200.ms_planner import DeviceBatch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8320
ctx = alore.Context(0)
prm = alore.default_params()
geom, grid = bench.build_world()
m = alore.SDFmap(ctx, gridmap_interval=0.05, detection_range=1e6, global_x_lower=geom.x_lower,
                 global_x_upper=geom.x_lower + (geom.glx - 0.5) * 0.05, global_y_lower=geom.y_lower,
                 global_y_upper=geom.y_lower + (geom.gly - 0.5) * 0.05)
m.gridmap_[:] = grid
m.has_map_ = True
m.forceUpdateESDF()
cands = bench.build_candidates(m.geom(), grid, m.distance_buffer_all_, 0, n)
res = {}
for mode in ("wave", "legacy"):
    if mode == "wave":
        os.environ["ALORE_OPT_WAVE"] = "1"
    else:
        os.environ.pop("ALORE_OPT_WAVE", None)
    db = DeviceBatch(ctx, cands)
    ms = []
    for it in range(3):
        db.run(prm)
        r = db.download()
        ms.append(db.kernel_ms())
    res[mode] = r
    print(mode, "kernel ms per run (first = cold order):", [round(x, 1) for x in ms])
    db.close()
a, b = res["wave"], res["legacy"]
print("bit-identical wave vs legacy:", np.array_equal(a.coeffs, b.coeffs), np.array_equal(a.cost, b.cost), np.array_equal(a.evals, b.evals))
