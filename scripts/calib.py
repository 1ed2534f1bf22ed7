import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import workloads, capi
from alore_legged_manipulator_b200.ms_planner import MSPlanner, DeviceBatch
from test_esdf_gpu import make_sdf
import oracle_lib
oracle_lib.load().orc_set_trig_portable(1)
ctx = alore.Context(0)
prm = alore.default_params()
n = 2048
t = time.time()
grid = workloads.random_map(n, n, 4, p_occ=0.0, p_unknown=0.0, wall=True, boxes=400, box_cells=(6, 30))
m = make_sdf(ctx, n, n, 0.05, grid); m.updateESDF2d()
pts = workloads.free_points(grid, m.geom(), m.distance_buffer_all_, 65, 4, min_clear=0.9)
print("map+esdf %.1fs" % (time.time() - t), flush=True)
t = time.time()
legs = workloads.leg_candidates(pts, headings=(0.0,), max_legs=int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
Ns = np.diff(legs.piece_off)
print("legs %d in %.1fs; pieces min/mean/max %d/%.1f/%d" % (legs.B, time.time() - t, Ns.min(), Ns.mean(), Ns.max()), flush=True)
db = DeviceBatch(ctx, legs)
for it in range(2):
    t = time.time(); db.run(prm); r = db.download(); dt = time.time() - t
    print("gpu run %d: %.3f s wall, kernel %.1f ms -> %.0f trajs/s; ok %d/%d evals mean %.0f max %d alm mean %.1f max %d replans mean %.2f" % (it, dt, db.kernel_ms(), legs.B / (db.kernel_ms() / 1e3), r.ok.sum(), legs.B, r.evals.mean(), r.evals.max(), r.alm_iters.mean(), r.alm_iters.max(), r.replans.mean()), flush=True)
sub = legs.subset(range(0, legs.B, max(1, legs.B // 64)))
t = time.time(); ref = oracle_lib.opt_batch(prm, m.geom(), m.distance_buffer_all_, sub, 8); dt = time.time() - t
print("cpu 8 threads: %d cands in %.2f s -> %.1f trajs/s (evals mean %.0f)" % (sub.B, dt, sub.B / dt, ref.evals.mean()))
