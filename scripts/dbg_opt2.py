import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import workloads, capi, front_end
from alore_legged_manipulator_b200.ms_planner import MSPlanner
from test_esdf_gpu import make_sdf
import oracle_lib
def rel(a, b): return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))
ctx = alore.Context(0)
oracle_lib.load().orc_set_trig_portable(1)
prm = alore.default_params(); prm.alm_max_outer = 20
geom, grid, cands = workloads.config1()
m = make_sdf(ctx, geom.glx, geom.gly, geom.grid_interval, grid); m.updateESDF2d()
pl = MSPlanner(ctx, prm, m); g = m.geom()
which = sys.argv[1]
if which == "pen":
    ref = oracle_lib.opt_batch(prm, g, m.distance_buffer_all_, cands, 1)
    N = cands.total_pieces
    po = np.array([0, N], np.int32)
    c, gC, gT, err = pl.penalty_batch(po, ref.coeffs, ref.piece_T, cands.start_xytheta[:, :2].copy(), cands.final_xytheta[:, :2].copy())
    cr, gCr, gTr, errr = oracle_lib.penalty_batch(prm, g, m.distance_buffer_all_, po, ref.coeffs, ref.piece_T, cands.start_xytheta[:, :2].copy(), cands.final_xytheta[:, :2].copy())
    print("penalty cost", c, cr, "gC rel", rel(gC, gCr), "gT rel", rel(gT, gTr), "err", err, errr)
else:
    n = int(which)
    # short straight leg with few pieces
    fe = front_end.FrontEndParams(); fe.mintrajNum = n
    ft = front_end.make_flat_traj([(-4.0, -4.0), (-4.0 + 0.2 * n, -4.0)], (-4.0, -4.0, 0.0), (-4.0 + 0.2 * n, -4.0, 0.0), fe)
    cb = front_end.pack_candidates([ft])
    print("N", cb.total_pieces)
    x0 = oracle_lib.initial_x(cb, 0)
    for stage in (0, 1):
        c, gr, err = pl.cost_batch(cb, stage, x0)
        cr, gref, eref = oracle_lib.cost(prm, g, m.distance_buffer_all_, cb, 0, stage, x0)
        print(f"stage{stage}: gpu {c[0]:.15g} ref {cr:.15g} grad rel {rel(gr, gref):.2e}")
