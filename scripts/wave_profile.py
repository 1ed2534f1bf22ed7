"""Per-round kernel timeline of the wavefront optimizer on the bench block (developer tool).
usage: python scripts/wave_profile.py [n_candidates] [out.csv]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import os
os.environ["ALORE_OPT_WAVE"] = "1"
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200.ms_planner import DeviceBatch, MSPlanner
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8320
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/wave_profile.csv"
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
print("warm kernel ms", db.kernel_ms())
t0 = time.perf_counter()
db.run(prm)
r = db.download()
print("second run kernel ms", db.kernel_ms(), "wall", time.perf_counter() - t0)
import ctypes as C
cnt = (C.c_ulonglong * 16)()
ctx.lib.alore_debug_wave_counters(ctx.h, cnt, 1)
db.run(prm)
db.download()
ctx.lib.alore_debug_wave_counters(ctx.h, cnt, 1)
c = list(cnt)
print("counters", c)
print(f"LU cycles/warp {c[0]/max(1,c[2]):.0f}  per pivot(kmax) {c[0]/max(1,c[12]):.0f}; back cycles/row {c[1]/max(1,c[12]):.0f}; reruns LU/back/adj {c[3]} {c[4]} {c[5]}")
print(f"adj upper cycles/row {c[6]/max(1,c[12]):.0f} lower {c[7]/max(1,c[12]):.0f}; two-loop cycles/history step {c[8]/max(1,c[9]):.0f}; step-kernel cycles/warp {c[10]/max(1,c[11]):.0f}")
os.environ["ALORE_WAVE_PROFILE"] = out
db.run(prm)
db.download()
del os.environ["ALORE_WAVE_PROFILE"]
a = np.loadtxt(out, delimiter=",", skiprows=1)
print("rounds", len(a), "sum ms solve/pen/adj/step", a[:, 2].sum(), a[:, 3].sum(), a[:, 4].sum(), a[:, 5].sum())
for lo, hi in ((0, 50), (50, 150), (150, 250), (250, 350), (350, 550), (550, 2000)):
    s = a[lo:hi]
    if len(s):
        print(f"rounds {lo}-{hi}: known~{s[:,1].mean():.0f} mean ms solve {s[:,2].mean():.4f} pen {s[:,3].mean():.4f} adj {s[:,4].mean():.4f} step {s[:,5].mean():.4f}")
