"""Per-candidate wall time inside the persistent optimizer kernel vs the same candidates alone (developer build:
make NVCCFLAGS+=-DALORE_CAND_TIMING into another .so, ALORE_B200_LIB=<it>).  Answers: how much slower does a heavy
candidate run when the SM is shared with 7 other candidates?   usage: python scripts/cand_timing.py [n]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200.ms_planner import DeviceBatch
from test_esdf_gpu import make_sdf
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8320
dump = str(ROOT / "gpurun_out" / "cand_ns.bin")
os.makedirs(ROOT / "gpurun_out", exist_ok=True)
os.environ["ALORE_CAND_TIMING_DUMP"] = dump
ctx = alore.Context(0)
prm = alore.default_params()
geom, grid = bench.build_world()
m = make_sdf(ctx, 2048, 2048, 0.05, grid)
m.updateESDF2d()
pts = bench.way_points(m.geom(), grid, m.distance_buffer_all_)
cands = bench.candidates_from_points(pts, 0, n)
N = np.diff(cands.piece_off)
db = DeviceBatch(ctx, cands)
for it in range(3):
    db.run(prm)
    r = db.download()
    print("run", it, "kernel ms", round(db.kernel_ms(), 1))
t = np.fromfile(dump, dtype=np.uint64).reshape(-1, 3)[:cands.B].astype(np.int64)
t0 = t[:, 0].min()
start, dur, sm = (t[:, 0] - t0) / 1e6, (t[:, 1] - t[:, 0]) / 1e6, t[:, 2]
ev = r.evals.astype(np.int64)
order = np.argsort(-dur)[:12]
print("makespan ms", (t[:, 1].max() - t0) / 1e6, " sum of durations / 1184 slots", dur.sum() / 1184)
print("heaviest candidates in the mix:   idx  N  evals  start_ms  dur_ms  ms/eval  sm")
for b in order:
    print(f"   {b:6d} {N[b]:3d} {ev[b]:6d} {start[b]:8.1f} {dur[b]:8.1f} {dur[b] / max(ev[b], 1):7.3f} {sm[b]:4d}")
solo = {}
for b in order[:4]:
    hb = DeviceBatch(ctx, cands.subset([int(b)]))
    hb.run(prm); hb.download()
    solo[int(b)] = hb.kernel_ms()
    hb.close()
print("same candidates alone on the GPU (kernel ms):", {k: round(v, 1) for k, v in solo.items()})
# ms/eval by piece count in the mix
for lo, hi in ((3, 20), (21, 40), (41, 60), (61, 80)):
    sel = (N >= lo) & (N <= hi) & (ev > 0)
    if sel.any():
        print(f"N {lo:2d}..{hi:2d}: {sel.sum():5d} candidates, mean ms/eval in the mix {np.mean(dur[sel] / ev[sel]):.3f}")
