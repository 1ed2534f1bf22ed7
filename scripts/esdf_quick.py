"""Quick ESDF timing on the GPU box (kernel-only via the library's CUDA events, and end to end)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import workloads
from test_esdf_gpu import make_sdf

ctx = alore.Context(0)
for name, (glx, gly), kw in [("bern0.02", (4096, 4096), dict(p_occ=0.02)), ("sparse1e-4", (4096, 4096), dict(p_occ=1e-4, wall=False)),
                             ("dense0.5", (4096, 4096), dict(p_occ=0.5)), ("single", (4096, 4096), None), ("8192x2048", (8192, 2048), dict(p_occ=0.02))]:
    if kw is None:
        grid = np.full(glx * gly, 1, np.uint8); grid[1000 * gly + 2000] = 2
    else:
        grid = workloads.random_map(glx, gly, 2, **kw)
    m = make_sdf(ctx, glx, gly, 0.05, grid)
    ts, ks = [], []
    for it in range(6):
        t0 = time.perf_counter(); m.updateESDF2d(); ts.append(time.perf_counter() - t0); ks.append(m.last_kernel_ms())
    cells = glx * gly
    print(f"{name:12s} kernels {min(ks[1:]):8.3f} ms  ({cells / min(ks[1:]) / 1e3:9.1f} Mcells/s, {13 * cells / min(ks[1:]) / 1e6:7.1f} GB/s algorithmic)  e2e {min(ts[1:]) * 1e3:8.2f} ms", flush=True)
