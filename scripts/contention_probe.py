"""What slows a candidate down when it shares the GPU?  Copies of ONE heavy candidate: B = warps_per_SM x SMs copies at
1, 2, 4, 8 resident warps per SM, with the reference's L-BFGS memory (256) and with a small one (8: no history streaming).
usage: python scripts/contention_probe.py [candidate index]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200.ms_planner import DeviceBatch
from test_esdf_gpu import make_sdf
import bench
idx = int(sys.argv[1]) if len(sys.argv) > 1 else 5110
ctx = alore.Context(0)
geom, grid = bench.build_world()
m = make_sdf(ctx, 2048, 2048, 0.05, grid)
m.updateESDF2d()
pts = bench.way_points(m.geom(), grid, m.distance_buffer_all_)
cands = bench.candidates_from_points(pts, 0, 8320)
nsm = 148
for mem in (256, 8):
    prm = alore.default_params()
    prm.lbfgs.mem_size = mem
    prm.path_lbfgs.mem_size = mem
    for wps in (1, 2, 4, 8):
        os.environ["ALORE_OPT_WARPS_PER_SM"] = str(wps)
        for copies in ((1, wps * nsm) if wps == 1 else (wps * nsm,)):
            db = DeviceBatch(ctx, cands.subset([idx] * copies))
            db.run(prm); r = db.download()
            db.run(prm); r = db.download()
            print(f"mem_size {mem:3d}  warps/SM {wps}  copies {copies:5d}: kernel ms {db.kernel_ms():8.1f}  evals/candidate {int(r.evals[0])}  ms/eval {db.kernel_ms() / max(int(r.evals[0]), 1):.3f}")
            db.close()
