"""GPU parity tests for the trajectory-optimizer path, through the C ABI, against the CPU oracle.

Tolerance (BASELINE north_star): penalties, gradients and optimised trajectories within 1e-9 relative.
Under the arithmetic contract of DESIGN.md (no FMA contraction, reference summation order, portable
sin/cos on both sides) the CUDA path is expected to be BIT-IDENTICAL to the oracle in `trig_portable`
mode; against the oracle's default glibc sin/cos a single evaluation agrees to ~1e-15 (asserted <= 1e-12),
while whole optimisations cannot be compared (the reference's optimizer amplifies 1e-15 to percent level;
see test_reference_optimizer_is_chaotic).
"""
import math

import numpy as np
import pytest

import oracle_lib
from alore_legged_manipulator_b200 import capi, front_end, workloads
from alore_legged_manipulator_b200.ms_planner import DeviceBatch, MSPlanner
from test_esdf_gpu import make_sdf

pytestmark = pytest.mark.gpu
REL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))) if a.size else 0.0


@pytest.fixture(scope="module")
def portable_trig():
    lib = oracle_lib.load()
    lib.orc_set_trig_portable(1)
    yield
    lib.orc_set_trig_portable(0)


@pytest.fixture(scope="module")
def world(ctx):
    """400x400 @0.05 m map with boxes + ESDF (GPU) + a planner with the reference's default parameters."""
    glx = gly = 400
    grid = workloads.random_map(glx, gly, 7, p_occ=0.0, p_unknown=0.0, wall=True, boxes=25, box_cells=(6, 24))
    m = make_sdf(ctx, glx, gly, 0.05, grid)
    m.updateESDF2d()
    prm = capi.default_params()
    prm.alm_max_outer = 20
    return m, prm, MSPlanner(ctx, prm, m), grid


def leg_batch(world, n_pts=9, max_legs=96, seed=5, **kw):
    m, prm, pl, grid = world
    pts = workloads.free_points(grid, m.geom(), m.distance_buffer_all_, n_pts, seed, min_clear=0.8)
    return workloads.leg_candidates(pts, headings=(0.0, 1.57), max_legs=max_legs, **kw)


# ---- single evaluations -----------------------------------------------------------------------------------
@pytest.mark.parametrize("stage", [0, 1])
def test_cost_and_gradient_batch(world, portable_trig, stage):
    m, prm, pl, grid = world
    cands = leg_batch(world, max_legs=64, dogleg=0.6)
    rng = np.random.default_rng(stage)
    x = np.concatenate([oracle_lib.initial_x(cands, b) for b in range(cands.B)])
    x = x + 0.03 * rng.normal(size=x.size)          # off the initial guess so every penalty family fires somewhere
    cost, g, err = pl.cost_batch(cands, stage, x)
    fired = 0
    for b in range(cands.B):
        o, n = cands.x_offset(b), 3 * int(cands.piece_off[b + 1] - cands.piece_off[b]) - 1
        cr, gr, er = oracle_lib.cost(prm, m.geom(), m.distance_buffer_all_, cands, b, stage, x[o:o + n])
        assert abs(cost[b] - cr) <= REL * abs(cr)
        assert rel(g[o:o + n], gr) <= REL
        if stage == 1:
            assert np.allclose(err[b], er, rtol=REL, atol=1e-15)
        assert cost[b] == cr and np.array_equal(g[o:o + n], gr)      # bit-identical under the contract
        fired += 1
    assert fired == cands.B


def test_cost_against_glibc_trig_oracle(world):
    """Against the oracle's DEFAULT (glibc sin/cos, what the reference calls) one evaluation agrees to ~1e-15."""
    m, prm, pl, grid = world
    cands = leg_batch(world, max_legs=32)
    x = np.concatenate([oracle_lib.initial_x(cands, b) for b in range(cands.B)])
    for stage in (0, 1):
        cost, g, _ = pl.cost_batch(cands, stage, x)
        for b in range(cands.B):
            o, n = cands.x_offset(b), 3 * int(cands.piece_off[b + 1] - cands.piece_off[b]) - 1
            cr, gr, _ = oracle_lib.cost(prm, m.geom(), m.distance_buffer_all_, cands, b, stage, x[o:o + n])
            assert abs(cost[b] - cr) <= 1e-12 * abs(cr) and rel(g[o:o + n], gr) <= 1e-12


def test_small_piece_counts_and_checkpoints(world, portable_trig, ctx):
    """N = 1, 2, 3 pieces; two collision check-points (plan_manager's car yaml) and the ICR (non-standard) model."""
    m, prm0, pl0, grid = world
    for variant in ("default", "two_checkpoints", "icr", "direct_v_omega"):
        prm = capi.default_params()
        if variant == "two_checkpoints":
            prm.n_checkpoints = 2
            prm.check_point[0][0], prm.check_point[0][1] = 0.3, 0.0
            prm.check_point[1][0], prm.check_point[1][1] = -0.3, 0.0
            prm.min_vel = 0.0
        elif variant == "icr":
            prm.if_standard_diff = 0
        elif variant == "direct_v_omega":
            prm.if_directly_constrain_v_omega = 1
        pl = MSPlanner(ctx, prm, m)
        fts = []
        for n in (1, 2, 3, 5):
            fe = front_end.FrontEndParams()
            fe.mintrajNum = n
            fe.timeResolution = 100.0     # piece count = mintrajNum
            fts.append(front_end.make_flat_traj([(-6.0, -6.0), (-6.0 + 0.25 * n, -5.8)], (-6.0, -6.0, 0.3),
                                                (-6.0 + 0.25 * n, -5.8, 1.0), fe))
        cands = front_end.pack_candidates(fts)
        assert list(np.diff(cands.piece_off)) == [1, 2, 3, 5]
        rng = np.random.default_rng(1)
        x = np.concatenate([oracle_lib.initial_x(cands, b) for b in range(cands.B)]) + 0.2 * rng.normal(size=cands.n_vars())
        for stage in (0, 1):
            cost, g, err = pl.cost_batch(cands, stage, x)
            for b in range(cands.B):
                o, n = cands.x_offset(b), 3 * int(cands.piece_off[b + 1] - cands.piece_off[b]) - 1
                cr, gr, _ = oracle_lib.cost(prm, m.geom(), m.distance_buffer_all_, cands, b, stage, x[o:o + n])
                assert cost[b] == cr and np.array_equal(g[o:o + n], gr), (variant, stage, b)


def test_norm_guard_quirk(world, portable_trig):
    """||x|| > 1e4  ->  cost 0 and g untouched (`#define inf 1 >> 30`, traj_representation.h:21)."""
    m, prm, pl, grid = world
    cands = leg_batch(world, max_legs=3)
    x = np.concatenate([oracle_lib.initial_x(cands, b) for b in range(cands.B)])
    x[cands.x_offset(1)] = 3e4
    g0 = np.full(x.size, 7.0)
    cost, g, _ = pl.cost_batch(cands, 1, x, g_init=g0)
    o, n = cands.x_offset(1), 3 * int(cands.piece_off[2] - cands.piece_off[1]) - 1
    assert cost[1] == 0.0 and np.all(g[o:o + n] == 7.0)
    assert cost[0] != 0.0 and not np.any(g[:o] == 7.0)


def test_penalty_batch_coefficient_space(world, portable_trig):
    """BASELINE config 3 shape (N = 64 pieces, K = 8) on a reduced batch, plus ragged piece counts."""
    m, prm, pl, grid = world
    po, coeffs, T, s_xy, f_xy = workloads.random_spline_batch(48, 64, m.geom(), m.distance_buffer_all_, grid, seed=3)
    cost, gC, gT, err = pl.penalty_batch(po, coeffs, T, s_xy, f_xy)
    cr, gCr, gTr, er = oracle_lib.penalty_batch(prm, m.geom(), m.distance_buffer_all_, po, coeffs, T, s_xy, f_xy, 4)
    assert np.all(np.abs(cost - cr) <= REL * np.abs(cr)) and rel(gC, gCr) <= REL and rel(gT, gTr) <= REL
    assert np.array_equal(cost, cr) and np.array_equal(gC, gCr) and np.array_equal(gT, gTr) and np.array_equal(err, er)
    # the collision branch must actually be exercised by this workload
    prm2 = capi.default_params()
    prm2.pw_collision = 0.0
    c2, _, _, _ = oracle_lib.penalty_batch(prm2, m.geom(), m.distance_buffer_all_, po, coeffs, T, s_xy, f_xy, 4)
    assert np.mean(c2 != cr) > 0.2
    # ragged: drop trailing pieces of every other trajectory
    keep, npo = [], [0]
    for b in range(8):
        n = 64 if b % 2 == 0 else 5 + b
        keep.extend(range(b * 64, b * 64 + n))
        npo.append(npo[-1] + n)
    npo = np.array(npo, np.int32)
    c3, gC3, gT3, e3 = pl.penalty_batch(npo, coeffs[keep], T[keep], s_xy[:8].copy(), f_xy[:8].copy())
    r3 = oracle_lib.penalty_batch(prm, m.geom(), m.distance_buffer_all_, npo, np.ascontiguousarray(coeffs[keep]),
                                  np.ascontiguousarray(T[keep]), s_xy[:8].copy(), f_xy[:8].copy(), 2)
    assert np.array_equal(c3, r3[0]) and np.array_equal(gC3, r3[1]) and np.array_equal(gT3, r3[2])


def test_final_collision_batch(world, portable_trig):
    m, prm, pl, grid = world
    po, coeffs, T, s_xy, f_xy = workloads.random_spline_batch(64, 24, m.geom(), m.distance_buffer_all_, grid, seed=9)
    col, md = pl.check_final_collision_batch(po, coeffs, T, s_xy)
    n_hit = 0
    for b in range(64):
        sl = slice(int(po[b]), int(po[b + 1]))
        c, d = oracle_lib.final_collision(prm, m.geom(), m.distance_buffer_all_, 24, coeffs[sl], T[sl], s_xy[b])
        assert col[b] == c and md[b] == d
        n_hit += c
    assert 0 < n_hit < 64


# ---- whole optimisations ------------------------------------------------------------------------------------
def check_results(res, ref, cands):
    assert np.array_equal(res.ok, ref.ok) and np.array_equal(res.status, ref.status)
    assert np.array_equal(res.replans, ref.replans) and np.array_equal(res.alm_iters, ref.alm_iters)
    # the oracle also counts the reference's printing evaluation after stage A (optimizer.cpp:341), once per optimizer() run
    assert np.array_equal(res.evals, ref.evals - ref.replans)
    worst = 0.0
    for b in range(cands.B):
        p0, p1 = int(cands.piece_off[b]), int(cands.piece_off[b + 1])
        worst = max(worst, rel(res.coeffs[p0:p1], ref.coeffs[p0:p1]), rel(res.piece_T[p0:p1], ref.piece_T[p0:p1]),
                    rel(res.inner_pts[p0 - b:p1 - b - 1], ref.inner_pts[p0 - b:p1 - b - 1]))
    assert worst <= REL, worst
    assert np.all(np.abs(res.cost - ref.cost) <= REL * np.abs(ref.cost)) and rel(res.tail_s, ref.tail_s) <= REL
    return worst


def test_minco_plan_config1(portable_trig):
    """BASELINE config 1: one trajectory on the 200x200 @0.05 m map, B = 1 through minco_plan()."""
    import alore_legged_manipulator_b200 as alore
    ctx = alore.Context(0)               # its own context: one context holds one device-resident map
    geom, grid, cands = workloads.config1()
    m = make_sdf(ctx, geom.glx, geom.gly, geom.grid_interval, grid)
    m.forceUpdateESDF()
    prm = capi.default_params()
    prm.alm_max_outer = 20
    pl = MSPlanner(ctx, prm, m)
    fts = [front_end.make_flat_traj([(-4.0, -4.0), (0.5, -0.5), (4.0, 4.0)], (-4.0, -4.0, 0.0), (4.0, 4.0, math.pi / 2))]
    assert pl.minco_plan(fts[0]) is True
    ref = oracle_lib.opt_batch(prm, m.geom(), m.distance_buffer_all_, cands, 1)
    worst = check_results(pl.last_result, ref, cands)
    assert worst == 0.0                                   # bit-identical
    dur, coef = pl.final_traj(0)
    assert dur.shape == (cands.total_pieces,) and coef.shape == (cands.total_pieces, 6, 2)
    assert pl.get_current_Innerpoints(0).shape == (2, cands.total_pieces - 1)
    assert pl.get_current_finState(0)[1, 0] == ref.tail_s[0]
    m.close()
    ctx.close()


def test_one_context_one_map():
    import alore_legged_manipulator_b200 as alore
    ctx = alore.Context(0)
    a = make_sdf(ctx, 64, 64, 0.1, workloads.random_map(64, 64, 1))
    b = make_sdf(ctx, 64, 64, 0.1, workloads.random_map(64, 64, 2))
    b.updateESDF2d()
    with pytest.raises(capi.AloreError):
        a.updateESDF2d()
    a.close()
    b.close()
    ctx.close()


def test_minco_plan_batch_of_legs(world, portable_trig):
    """Ragged batch (N = 6..27) of task-planner legs: every optimised trajectory bit-identical to the oracle."""
    m, prm, pl, grid = world
    cands = leg_batch(world, max_legs=96)
    res = pl.minco_plan_batch(cands)
    ref = oracle_lib.opt_batch(prm, m.geom(), m.distance_buffer_all_, cands, 8)
    assert check_results(res, ref, cands) == 0.0
    assert res.ok.sum() > 0.8 * cands.B


def test_collision_replans_and_cut_trajectories(world, portable_trig, ctx):
    """Forces the final-collision retry path (time_weight *= 0.75) and the `if_cut` ALM parameter set."""
    m, prm0, pl0, grid = world
    prm = capi.default_params()
    prm.alm_max_outer = 6
    prm.finalMinSafeDis = 0.45         # demanding clearance -> some candidates replan or fail
    prm.safeReplanMaxTime = 2
    pl = MSPlanner(ctx, prm, m)
    fe = front_end.FrontEndParams()
    fe.trajCutLength = 4.0             # long legs are cut: if_cut = true
    pts = workloads.free_points(grid, m.geom(), m.distance_buffer_all_, 7, 11, min_clear=0.5)
    cands = workloads.leg_candidates(pts, headings=(0.0,), fe=fe, max_legs=40)
    assert cands.if_cut.sum() > 0
    res = pl.minco_plan_batch(cands)
    ref = oracle_lib.opt_batch(prm, m.geom(), m.distance_buffer_all_, cands, 8)
    check_results(res, ref, cands)
    assert res.replans.max() == 2 and (res.ok == 0).sum() > 0 and (res.ok == 1).sum() > 0


def test_resident_batch_api_and_argmin(world, portable_trig):
    m, prm, pl, grid = world
    cands = leg_batch(world, max_legs=40)
    db = DeviceBatch(pl.ctx, cands)
    db.run(prm)
    r1 = db.download()
    db.run(prm)                          # re-running a resident batch is deterministic
    r2 = db.download()
    assert np.array_equal(r1.coeffs, r2.coeffs) and np.array_equal(r1.cost, r2.cost)
    assert db.kernel_ms() > 0.0
    bc, bi = db.argmin()
    ok = np.flatnonzero(r1.ok == 1)
    assert bi == ok[np.argmin(r1.cost[ok])] and bc == r1.cost[bi]
    db.close()


def test_result_independent_of_batch_composition(world, portable_trig):
    """A candidate's result does not depend on which other candidates share the launch (sharding invariance)."""
    m, prm, pl, grid = world
    cands = leg_batch(world, max_legs=24)
    full = pl.minco_plan_batch(cands)
    for idx in ([5], [0, 7, 13], list(range(23, 11, -1))):
        sub = cands.subset(idx)
        r = pl.minco_plan_batch(sub)
        for k, b in enumerate(idx):
            p0, p1 = int(cands.piece_off[b]), int(cands.piece_off[b + 1])
            q0, q1 = int(sub.piece_off[k]), int(sub.piece_off[k + 1])
            assert np.array_equal(r.coeffs[q0:q1], full.coeffs[p0:p1]) and r.cost[k] == full.cost[b]


def test_reference_optimizer_is_chaotic(world):
    """Why trajectory-level parity needs identical arithmetic: the ORACLE's own result moves by far more than
    1e-9 when one input is perturbed by 1e-15 relative.  (CPU only; documents the claim in DESIGN.md.)"""
    m, prm, pl, grid = world
    cands = leg_batch(world, max_legs=12)
    ref = oracle_lib.opt_batch(prm, m.geom(), m.distance_buffer_all_, cands, 4)
    pert = capi.CandidateBatch.concat([cands])
    pert.inner_pts[:, 0] *= (1 + 1e-15)
    r2 = oracle_lib.opt_batch(prm, m.geom(), m.distance_buffer_all_, pert, 4)
    rels = [rel(r2.coeffs[int(cands.piece_off[b]):int(cands.piece_off[b + 1])],
                ref.coeffs[int(cands.piece_off[b]):int(cands.piece_off[b + 1])]) for b in range(cands.B)]
    assert max(rels) > 1e-6


def test_missing_map_is_an_error(ctx):
    import alore_legged_manipulator_b200 as alore
    c2 = alore.Context(0)
    geom, grid, cands = workloads.config1()
    res = capi.ResultBatch(cands)
    cs, rs = cands.as_struct(), res.as_struct()
    import ctypes as C
    prm = capi.default_params()
    rc = c2.lib.alore_opt_batch(c2.h, C.byref(prm), C.byref(cs), C.byref(rs))
    assert rc == -3 and b"ESDF" in c2.lib.alore_last_error(c2.h)
    c2.close()


def test_split_division_is_ieee_division(ctx):
    """rcp_refine + div_rcp (csrc/traj_opt.cuh) replace `a / b` on the dependent chains of the banded LU and of the
    triangular sweeps (minco.hpp:99-197).  They must return the bits of the compiler's IEEE division for every
    operand pair: raw 64-bit patterns (all exponents, subnormals, Inf/NaN), solver-range magnitudes, near-equal
    mantissas.  2^26 pairs per seed."""
    import ctypes as C
    for seed in (1, 0xA105E, 2 ** 40 + 7):
        bad = C.c_longlong(-1)
        ctx.check(ctx.lib.alore_selftest_division(ctx.h, 1 << 26, seed, C.byref(bad)))
        assert bad.value == 0, f"seed {seed}: {bad.value} quotients differ from a / b"


def test_exact_division_rerun_path_gives_same_bits(world, portable_trig, ctx):
    """The banded-solver passes run with the speculative split division and are repeated with the compiler's
    division when a quotient leaves the fast path's range.  Forcing that second path must not change a single bit."""
    m, prm, pl, grid = world
    legs = leg_batch(world, max_legs=24)
    a = pl.minco_plan_batch(legs)
    ctx.check(ctx.lib.alore_debug_force_exact_division(ctx.h, 1))
    try:
        b = pl.minco_plan_batch(legs)
    finally:
        ctx.check(ctx.lib.alore_debug_force_exact_division(ctx.h, 0))
    assert np.array_equal(a.coeffs, b.coeffs) and np.array_equal(a.cost, b.cost) and np.array_equal(a.evals, b.evals)
    assert np.array_equal(a.piece_T, b.piece_T) and np.array_equal(a.status, b.status)


def test_rescheduled_second_tick_gives_same_bits(world, portable_trig, ctx):
    """A resident batch that is optimised again is handed out longest-predicted-work-first (evaluation counts of the
    previous tick).  Only the queue order changes: every result of the second tick must equal the first bit for bit,
    and the one-shot API (which learns from the context's memory of the same batch structure) must agree as well."""
    m, prm, pl, grid = world
    cands = leg_batch(world, max_legs=64)
    db = DeviceBatch(ctx, cands)
    db.run(prm)
    a = db.download()
    db.run(prm)
    b = db.download()
    db.close()
    c = pl.minco_plan_batch(cands)
    for r in (b, c):
        assert np.array_equal(a.coeffs, r.coeffs) and np.array_equal(a.cost, r.cost) and np.array_equal(a.evals, r.evals)
        assert np.array_equal(a.piece_T, r.piece_T) and np.array_equal(a.status, r.status) and np.array_equal(a.ok, r.ok)


def test_long_trajectories_and_fine_sampling(portable_trig):
    """Long corridor legs (N ~ 50..130 pieces: the 8-elements-per-lane two-loop window, multi-chunk sweeps, the larger
    shared-memory carve-up) and sparseResolution = 12 (cell-prefix chunking, larger term log) stay bit-identical."""
    import alore_legged_manipulator_b200 as alore
    ctx = alore.Context(0)
    glx, gly = 3200, 160                                    # 160 m x 8 m at 0.05 m
    grid = workloads.random_map(glx, gly, 11, p_occ=0.0, p_unknown=0.0, wall=True, boxes=30, box_cells=(4, 12))
    grid.reshape(glx, gly)[:, 60:100] = capi.UNOCCUPIED     # a free lane down the middle
    m = make_sdf(ctx, glx, gly, 0.05, grid)
    m.updateESDF2d()
    fts = []
    for x1, th in ((-10.0, 0.0), (40.0, 0.3), (75.0, 0.0)):
        fts.append(front_end.make_flat_traj([(-75.0, 0.0), (x1, 0.0)], (-75.0, 0.0, 0.0), (x1, 0.0, th)))
    cands = front_end.pack_candidates(fts)
    assert int(np.diff(cands.piece_off).max()) > 100
    for K in (8, 12):
        prm = capi.default_params()
        prm.alm_max_outer = 6
        prm.sparseResolution = K
        pl = MSPlanner(ctx, prm, m)
        res = pl.minco_plan_batch(cands)
        ref = oracle_lib.opt_batch(prm, m.geom(), m.distance_buffer_all_, cands, 4)
        assert check_results(res, ref, cands) == 0.0
    m.close()
    ctx.close()
