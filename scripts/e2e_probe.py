"""Where does the end-to-end (host-buffer C ABI) step spend its time?  python scripts/e2e_probe.py"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ctypes as C
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import capi
from alore_legged_manipulator_b200.ms_planner import MSPlanner, DeviceBatch
import bench
from test_esdf_gpu import make_sdf

ctx = alore.Context(0)
prm = alore.default_params()
geom, grid = bench.build_world()
m = make_sdf(ctx, geom.glx, geom.gly, 0.05, grid)
m.updateESDF2d()
cands = bench.build_candidates(m.geom(), grid, m.distance_buffer_all_, 0, bench.PER_GPU)
pl = MSPlanner(ctx, prm, m)
for it in range(3):
    t0 = time.perf_counter(); m.updateESDF2d(); t1 = time.perf_counter()
    res = capi.ResultBatch(cands); cs, rs = cands.as_struct(), res.as_struct(); t2 = time.perf_counter()
    bh = C.c_void_p()
    ctx.check(ctx.lib.alore_batch_upload(ctx.h, C.byref(cs), C.byref(bh))); t3 = time.perf_counter()
    ctx.check(ctx.lib.alore_batch_run(ctx.h, C.byref(prm), bh, None)); t4 = time.perf_counter()
    ctx.check(ctx.lib.alore_batch_download(ctx.h, bh, C.byref(rs))); t5 = time.perf_counter()
    ms = C.c_float(); ctx.lib.alore_batch_last_kernel_ms(bh, C.byref(ms))
    ctx.lib.alore_batch_free(bh); t6 = time.perf_counter()
    t7 = time.perf_counter(); r = pl.minco_plan_batch(cands); t8 = time.perf_counter()
    print(f"iter {it}: esdf {1e3*(t1-t0):.1f} ms | alloc results {1e3*(t2-t1):.1f} | upload {1e3*(t3-t2):.1f} | launch {1e3*(t4-t3):.1f} | "
          f"wait+download {1e3*(t5-t4):.1f} (kernel {ms.value:.1f}) | free {1e3*(t6-t5):.1f} || alore_opt_batch total {1e3*(t8-t7):.1f}")
