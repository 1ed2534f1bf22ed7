"""One short wavefront optimisation for an ncu capture (developer tool): python scripts/wave_round_for_ncu.py [n]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import os
os.environ["ALORE_OPT_WAVE"] = "1"
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200.ms_planner import DeviceBatch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4160
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
db = DeviceBatch(ctx, cands)
db.run(prm)
db.download()
