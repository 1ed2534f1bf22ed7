import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import ctypes as C, numpy as np, torch
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import workloads
import bench
from test_esdf_gpu import make_sdf
ctx = alore.Context(0); prm = alore.default_params()
geom, grid = bench.build_world()
m = make_sdf(ctx, geom.glx, geom.gly, 0.05, grid); m.updateESDF2d(); gm=m.geom()
Bp, Np = 4096, 64
po, coeffs, Tp, s_xy, f_xy = workloads.random_spline_batch(Bp, Np, gm, m.distance_buffer_all_, grid, seed=3)
dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
d_po, d_c, d_T, d_s, d_f = dev(po, np.int32), dev(coeffs, np.float64), dev(Tp, np.float64), dev(s_xy, np.float64), dev(f_xy, np.float64)
d_cost = torch.zeros(Bp, dtype=torch.float64, device="cuda"); d_gC = torch.zeros(Bp*Np*12, dtype=torch.float64, device="cuda")
d_gT = torch.zeros(Bp*Np, dtype=torch.float64, device="cuda"); d_err = torch.zeros(Bp*2, dtype=torch.float64, device="cuda")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sptr = C.c_void_p(stream.cuda_stream)
pv = lambda t: C.c_void_p(t.data_ptr())
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
for a, b in evs:
    a.record(stream)
    ctx.check(ctx.lib.alore_penalty_batch_dev(ctx.h, C.byref(prm), Bp, Np, pv(d_po), pv(d_c), pv(d_T), pv(d_s), pv(d_f), pv(d_cost), pv(d_gC), pv(d_gT), pv(d_err), sptr))
    b.record(stream)
torch.cuda.synchronize()
print("penalty ms", sorted(a.elapsed_time(b) for a, b in evs)[1:5])
