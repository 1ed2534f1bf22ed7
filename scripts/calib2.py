import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import workloads, capi
from alore_legged_manipulator_b200.ms_planner import MSPlanner, DeviceBatch
from test_esdf_gpu import make_sdf
import bench
ctx = alore.Context(0)
geom, grid = bench.build_world()
m = make_sdf(ctx, geom.glx, geom.gly, 0.05, grid); m.updateESDF2d()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8320
cands = bench.build_candidates(m.geom(), grid, m.distance_buffer_all_, 0, B)
db = DeviceBatch(ctx, cands)
for mem in [int(a) for a in sys.argv[2:]] or [256]:
    prm = alore.default_params(); prm.lbfgs.mem_size = mem; prm.path_lbfgs.mem_size = mem
    db.run(prm); r = db.download()
    ab, ev, it = db.stats()
    print(f"mem_size {mem}: kernel {db.kernel_ms():.1f} ms -> {B / db.kernel_ms() * 1e3:.0f} trajs/s; evals {ev} iters {it} ok {int(r.ok.sum())} alg GB {ab/1e9:.0f}", flush=True)
