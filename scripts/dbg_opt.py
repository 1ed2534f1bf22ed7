import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import workloads, capi
from alore_legged_manipulator_b200.ms_planner import MSPlanner, DeviceBatch
from test_esdf_gpu import make_sdf
import oracle_lib

def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))

ctx = alore.Context(0)
oracle_lib.load().orc_set_trig_portable(1)
prm = alore.default_params(); prm.alm_max_outer = 20
# ---- config 1
geom, grid, cands = workloads.config1()
m = make_sdf(ctx, geom.glx, geom.gly, geom.grid_interval, grid)
m.updateESDF2d()
pl = MSPlanner(ctx, prm, m)
g = m.geom()
x0 = oracle_lib.initial_x(cands, 0)
for stage in (0, 1):
    c, gr, err = pl.cost_batch(cands, stage, x0)
    cr, gref, eref = oracle_lib.cost(prm, g, m.distance_buffer_all_, cands, 0, stage, x0)
    print(f"cfg1 cost stage{stage}: gpu {c[0]:.15g} ref {cr:.15g} rel {abs(c[0]-cr)/abs(cr):.2e} grad rel {rel(gr, gref):.2e} err {err[0]} {eref}")
t = time.time(); res = pl.minco_plan_batch(cands); t = time.time() - t
ref = oracle_lib.opt_batch(prm, g, m.distance_buffer_all_, cands, 1)
print(f"cfg1 plan: gpu ok {res.ok[0]} st {res.status[0]} rp {res.replans[0]} alm {res.alm_iters[0]} ev {res.evals[0]} cost {res.cost[0]:.12g}  ({t*1e3:.1f} ms)")
print(f"           ref ok {ref.ok[0]} st {ref.status[0]} rp {ref.replans[0]} alm {ref.alm_iters[0]} ev {ref.evals[0]} cost {ref.cost[0]:.12g}")
print(f"           coeff rel {rel(res.coeffs, ref.coeffs):.3e}  T rel {rel(res.piece_T, ref.piece_T):.3e}")
# ---- batch of legs on a 400x400 map
glx = gly = 400
gi = 0.05
grid = workloads.random_map(glx, gly, 7, p_occ=0.0, p_unknown=0.0, wall=True, boxes=25, box_cells=(6, 24))
m2 = make_sdf(ctx, glx, gly, gi, grid)
m2.updateESDF2d()
g2 = m2.geom()
pts = workloads.free_points(grid, g2, m2.distance_buffer_all_, 9, 5, min_clear=0.8)
legs = workloads.leg_candidates(pts, headings=(0.0, 1.57), max_legs=96)
print("legs", legs.B, "pieces min/max", np.diff(legs.piece_off).min(), np.diff(legs.piece_off).max())
pl2 = MSPlanner(ctx, prm, m2)
t = time.time(); res = pl2.minco_plan_batch(legs); tg = time.time() - t
t = time.time(); ref = oracle_lib.opt_batch(prm, g2, m2.distance_buffer_all_, legs, 8); tc = time.time() - t
print(f"batch: gpu {tg*1e3:.1f} ms, cpu(8 thr) {tc*1e3:.1f} ms; ok gpu {res.ok.sum()} ref {ref.ok.sum()}; evals gpu {res.evals.sum()} ref {ref.evals.sum()}")
rels = []
for b in range(legs.B):
    p0, p1 = legs.piece_off[b], legs.piece_off[b + 1]
    rels.append(rel(res.coeffs[p0:p1], ref.coeffs[p0:p1]))
rels = np.array(rels)
print("coeff rel diff per candidate: median %.2e, <=1e-9: %d/%d, <=1e-6: %d, max %.2e" % (np.median(rels), (rels <= 1e-9).sum(), legs.B, (rels <= 1e-6).sum(), rels.max()))
print("same evals:", int((res.evals == ref.evals).sum()), "same status:", int((res.status == ref.status).sum()), "same ok", int((res.ok == ref.ok).sum()))
bad = np.argsort(-rels)[:5]
for b in bad:
    print("  cand", b, "N", legs.piece_off[b+1]-legs.piece_off[b], "rel", rels[b], "evals", res.evals[b], ref.evals[b], "cost", res.cost[b], ref.cost[b], "alm", res.alm_iters[b], ref.alm_iters[b], "rp", res.replans[b], ref.replans[b])
