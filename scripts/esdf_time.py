"""ESDF kernel time on the bench maps: python scripts/esdf_time.py [bernoulli|corridor|boxes]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import workloads
from test_esdf_gpu import make_sdf
ctx = alore.Context(0)
for which in sys.argv[1:] or ["bernoulli", "corridor", "boxes"]:
    if which == "bernoulli":
        nx, ny, gi = 4096, 4096, 0.05
        grid = workloads.random_map(nx, ny, 2, p_occ=0.02, p_unknown=0.01)
    elif which == "corridor":
        nx, ny, gi = 8192, 2048, 0.005
        grid = workloads.corridor_map(nx, ny, 5, width_cells=400, clutter=0.02)
    else:
        nx, ny, gi = 2048, 2048, 0.05
        grid = workloads.random_map(nx, ny, 4, p_occ=0.0, p_unknown=0.0, wall=True, boxes=400, box_cells=(6, 30))
    m = make_sdf(ctx, nx, ny, gi, grid)
    ts = []
    for _ in range(5):
        m.updateESDF2d()
        ts.append(m.last_kernel_ms())
    print(f"{which:10s} {nx}x{ny}: kernels ms {min(ts):.4f}  ({nx*ny/min(ts)/1e3:.0f} Mcells/s)")
    m.close()
