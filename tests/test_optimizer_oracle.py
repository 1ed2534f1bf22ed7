"""CPU tests pinning the optimizer part of the oracle with independent known-answer checks
(the reference ships no tests for this path): MINCO invariants, adjoint vs dense transpose solve,
finite-difference gradients, Rosenbrock L-BFGS, a full minco_plan on BASELINE config 1, and the
host-side front-end restatement against the oracle's."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_lib
from alore_legged_manipulator_b200 import capi, front_end, workloads


def default_params(**kw) -> capi.Params:
    p = capi.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.fixture(scope="module")
def cfg1():
    geom, grid, cands = workloads.config1()
    dist = np.full(geom.glx * geom.gly, np.finfo(np.float64).max)
    oracle_lib.esdf_update(geom, grid, (0, 0), (geom.glx - 1, geom.gly - 1), dist)
    return geom, grid, cands, dist


# ---- MINCO -------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1, 2, 3, 7, 20])
def test_minco_solution_satisfies_system_and_continuity(N):
    rng = np.random.default_rng(N)
    head, tail = rng.normal(size=(2, 3)), rng.normal(size=(2, 3))
    inPs = rng.normal(size=(max(N - 1, 0), 2))
    T = rng.uniform(0.3, 0.9, size=N)
    c, energy, gdC, gdT = oracle_lib.minco_solve(N, head, tail, inPs, T, (0.33, 1.0))
    A = np.zeros((6 * N, 6 * N))
    oracle_lib.load().orc_minco_matrix(N, capi.dptr(T), capi.dptr(A))
    b = np.zeros((6 * N, 2))
    b[0:3] = head.T
    for i in range(N - 1):
        b[6 * i + 5] = inPs[i]
    b[6 * N - 3:] = tail.T
    assert np.allclose(A @ c, b, rtol=0, atol=1e-9 * max(1.0, np.abs(c).max()))
    # agrees with an independent banded solve (scipy) of the same system
    c2 = workloads.minco_coeffs(head, tail, inPs, T)
    assert np.allclose(c, c2, rtol=1e-9, atol=1e-9)

    def ev(i, t, der):
        k = np.arange(6)
        if der == 0:
            basis = t ** k
        else:
            basis = np.zeros(6)
            for kk in range(der, 6):
                basis[kk] = math.perm(kk, der) * t ** (kk - der)
        return basis @ c[6 * i:6 * i + 6]
    # head / tail PVA, pass through inner points, C4 continuity at knots
    for d in range(3):
        assert np.allclose(ev(0, 0.0, d), head[:, d], atol=1e-9)
        assert np.allclose(ev(N - 1, T[-1], d), tail[:, d], atol=1e-8)
    for i in range(N - 1):
        assert np.allclose(ev(i, T[i], 0), inPs[i], atol=1e-9)
        for d in range(5):
            assert np.allclose(ev(i, T[i], d), ev(i + 1, 0.0, d), atol=1e-7 * max(1.0, abs(ev(i, T[i], d)).max()))
    # energy = integral of weighted squared jerk (Gauss-Legendre check)
    xs, ws = np.polynomial.legendre.leggauss(8)
    e = 0.0
    for i in range(N):
        t = 0.5 * T[i] * (xs + 1)
        for tt, w in zip(t, ws):
            j = ev(i, tt, 3)
            e += 0.5 * T[i] * w * (0.33 * j[0] ** 2 + 1.0 * j[1] ** 2)
    assert math.isclose(e, energy, rel_tol=1e-9)


def test_minco_adjoint_matches_dense_transpose_solve():
    N = 6
    rng = np.random.default_rng(1)
    head, tail = rng.normal(size=(2, 3)), rng.normal(size=(2, 3))
    inPs = rng.normal(size=(N - 1, 2))
    T = rng.uniform(0.3, 0.9, size=N)
    pgC, pgT = rng.normal(size=(6 * N, 2)), rng.normal(size=N)
    gradP, gradT, gradTail = np.zeros((N - 1, 2)), np.zeros(N), np.zeros(2)
    ew = np.array([0.33, 1.0])
    oracle_lib.load().orc_minco_adjoint(N, capi.dptr(np.ascontiguousarray(head)), capi.dptr(np.ascontiguousarray(tail)),
                                        capi.dptr(np.ascontiguousarray(inPs)), capi.dptr(T), capi.dptr(ew),
                                        capi.dptr(pgC), capi.dptr(pgT), capi.dptr(gradP), capi.dptr(gradT), capi.dptr(gradTail))
    A = np.zeros((6 * N, 6 * N))
    oracle_lib.load().orc_minco_matrix(N, capi.dptr(T), capi.dptr(A))
    adj = np.linalg.solve(A.T, pgC)
    for i in range(N - 1):
        assert np.allclose(gradP[i], adj[6 * i + 5], rtol=1e-8, atol=1e-8)
    assert np.allclose(gradTail, adj[6 * N - 3], rtol=1e-8, atol=1e-8)
    # gradT: finite differences of  L(T) = <pgC, c(T)> + <pgT, T>  (c = MINCO solution)
    def L(Tv):
        c, _, _, _ = oracle_lib.minco_solve(N, head, tail, inPs, Tv, ew)
        return float((pgC * c).sum())
    for i in range(N):
        h = 1e-6
        Tp, Tm = T.copy(), T.copy()
        Tp[i] += h
        Tm[i] -= h
        fd = (L(Tp) - L(Tm)) / (2 * h) + pgT[i]
        assert math.isclose(fd, gradT[i], rel_tol=2e-5, abs_tol=1e-5)


# ---- cost gradients vs finite differences ---------------------------------------------------------
def _fd_check(prm, geom, dist, cands, stage, x0, sl, rng, rel):
    f, g, _ = oracle_lib.cost(prm, geom, dist, cands, 0, stage, x0)
    v = np.zeros_like(x0)
    v[sl] = rng.normal(size=v[sl].size)
    v /= np.linalg.norm(v)
    h = 1e-6
    fp, _, _ = oracle_lib.cost(prm, geom, dist, cands, 0, stage, x0 + h * v)
    fm, _, _ = oracle_lib.cost(prm, geom, dist, cands, 0, stage, x0 - h * v)
    fd = (fp - fm) / (2 * h)
    assert math.isclose(fd, g @ v, rel_tol=rel, abs_tol=1e-4 * max(1.0, abs(f)) * 1e-3), (fd, g @ v)


def test_cost_gradient_finite_differences(cfg1):
    geom, grid, cands, dist = cfg1
    N = cands.total_pieces
    rng = np.random.default_rng(0)
    x0 = oracle_lib.initial_x(cands, 0) + 0.05 * rng.normal(size=3 * N - 1)
    Psl, tail, tau = slice(0, 2 * (N - 1)), slice(2 * (N - 1), 2 * (N - 1) + 1), slice(2 * (N - 1) + 1, 3 * N - 1)
    lib = oracle_lib.load()
    try:
        # stage 1, every term, with the exact own-sample chain weight (test hook): analytic == FD
        lib.orc_set_exact_chain_weights(1)
        for sl in (Psl, tail, tau):
            _fd_check(default_params(), geom, dist, cands, 1, x0, sl, rng, 1e-6)
    finally:
        lib.orc_set_exact_chain_weights(0)
    # reference behaviour: exact without the collision term ...
    for sl in (Psl, tail, tau):
        _fd_check(default_params(pw_collision=0.0), geom, dist, cands, 1, x0, sl, rng, 1e-6)
    # ... and the collision term's gradient is only approximate BY DESIGN (documented in the oracle header)
    f, g, _ = oracle_lib.cost(default_params(), geom, dist, cands, 0, 1, x0)
    assert np.isfinite(f) and np.all(np.isfinite(g))
    # stage 0: inner points and tail exact; tau differs by the reference's weight mix-up
    # (cost uses PathpenaltyWt.time_weight = 20, gradient penaltyWt.time_weight = 50: optimizer.cpp:1308 vs 1312)
    prm = default_params()
    for sl in (Psl, tail):
        _fd_check(prm, geom, dist, cands, 0, x0, sl, rng, 1e-6)
    f, g, _ = oracle_lib.cost(prm, geom, dist, cands, 0, 0, x0)
    h = 1e-6
    v = np.zeros_like(x0)
    v[tau] = rng.normal(size=N)
    fp, _, _ = oracle_lib.cost(prm, geom, dist, cands, 0, 0, x0 + h * v)
    fm, _, _ = oracle_lib.cost(prm, geom, dist, cands, 0, 0, x0 - h * v)
    t = x0[tau]
    dT = np.where(t > 0, t + 1.0, (1.0 - t) / ((0.5 * t - 1.0) * t + 1.0) ** 2)
    quirk = (prm.pw_time - prm.ppw_time) * float(dT @ v[tau])
    assert math.isclose((fp - fm) / (2 * h) + quirk, g @ v, rel_tol=1e-6)


def test_cost_norm_guard_returns_zero_and_keeps_gradient(cfg1):
    geom, grid, cands, dist = cfg1
    x = oracle_lib.initial_x(cands, 0)
    x[0] = 2e4
    lib = oracle_lib.load()
    g = np.full(x.size, 7.0)
    c = C.c_double(-1.0)
    cs = cands.as_struct()
    prm = default_params()
    lib.orc_cost(C.byref(prm), C.byref(geom), capi.dptr(dist), C.byref(cs), 0, 1, capi.dptr(x), None, None, 0.6, C.byref(c),
                 capi.dptr(g), None)
    assert c.value == 0.0 and np.all(g == 7.0)   # `#define inf 1 >> 30` (traj_representation.h:21)


# ---- L-BFGS ------------------------------------------------------------------------------------------
def test_lbfgs_rosenbrock_known_answer():
    lib = oracle_lib.load()
    n = 20
    x = np.tile([-1.2, 1.0], n // 2).astype(np.float64)
    prm = capi.default_params().lbfgs
    prm.mem_size, prm.past, prm.delta, prm.g_epsilon, prm.max_iterations = 8, 0, 0.0, 1e-8, 2000
    f, it, ev = C.c_double(), (C.c_int32 * 1)(), (C.c_int32 * 1)()
    ret = lib.orc_lbfgs_rosenbrock(n, capi.dptr(x), C.byref(prm), C.byref(f), it, ev)
    assert ret == 0
    assert np.allclose(x, 1.0, atol=1e-6) and f.value < 1e-12


def test_lbfgs_early_accept_quirk_changes_behaviour():
    """The ALORE-specific early return (lbfgs.hpp:326-329) fires only when past > 0."""
    lib = oracle_lib.load()
    n = 10
    res = {}
    for past in (0, 3):
        x = np.tile([-1.2, 1.0], n // 2).astype(np.float64)
        prm = capi.default_params().lbfgs
        prm.mem_size, prm.past, prm.delta, prm.g_epsilon, prm.max_iterations = 8, past, 5e-4, 0.0, 500
        f, it, ev = C.c_double(), (C.c_int32 * 1)(), (C.c_int32 * 1)()
        res[past] = (lib.orc_lbfgs_rosenbrock(n, capi.dptr(x), C.byref(prm), C.byref(f), it, ev), it[0], f.value)
    assert res[3][0] == 1            # LBFGS_STOP through the past/delta test
    assert res[0][0] != 1            # without `past` only max_iterations or a line-search error can end it


# ---- full minco_plan ------------------------------------------------------------------------------------
def test_minco_plan_config1(cfg1):
    geom, grid, cands, dist = cfg1
    prm = default_params(alm_max_outer=20)
    res = oracle_lib.opt_batch(prm, geom, dist, cands, 1)
    assert res.ok[0] == 1 and res.replans[0] >= 1 and res.evals[0] > 10
    N = cands.total_pieces
    T = res.piece_T
    assert np.all(T > 0.05) and np.all(T < 5.0)
    # the optimised coefficients are a valid MINCO spline for the optimised points / times
    head, tail = cands.start_state[0], cands.final_state[0].copy()
    tail[1, 0] = res.tail_s[0]
    c, _, _, _ = oracle_lib.minco_solve(N, head, tail, res.inner_pts[: N - 1], T, (prm.energyWeights[0], prm.energyWeights[1]))
    assert np.allclose(c, res.coeffs.reshape(6 * N, 2), rtol=1e-9, atol=1e-9)
    # end-point equality constraint met, final trajectory collision-free at finalMinSafeDis
    col, md = oracle_lib.final_collision(prm, geom, dist, N, res.coeffs, T, cands.start_xytheta[0, :2])
    assert col == 0 and md >= prm.finalMinSafeDis
    # determinism, and thread-count invariance of the batch runner
    both = capi.CandidateBatch.concat([cands, cands, cands])
    r3 = oracle_lib.opt_batch(prm, geom, dist, both, 3)
    for k in range(3):
        assert np.array_equal(r3.coeffs[k * N:(k + 1) * N], res.coeffs)


# ---- front end (host logic) ----------------------------------------------------------------------------
def test_front_end_python_matches_oracle_restatement():
    lib = oracle_lib.load()
    rng = np.random.default_rng(3)
    fe = front_end.FrontEndParams()
    for trial in range(20):
        npts = int(rng.integers(2, 6))
        path = rng.uniform(-8, 8, size=(npts, 2))
        start = (path[0, 0], path[0, 1], float(rng.uniform(-3, 3)))
        end = (path[-1, 0], path[-1, 1], float(rng.uniform(-3, 3)))
        fe.trajCutLength = 600.0 if trial % 3 else 6.0
        ft = front_end.make_flat_traj([tuple(p) for p in path], start, end, fe)
        cap = 512
        inner, initT, pos = np.zeros((cap, 2)), C.c_double(), np.zeros((cap, 3))
        ss, fs, fx, cut = np.zeros((2, 3)), np.zeros((2, 3)), np.zeros(3), np.zeros(1, np.uint8)
        z3 = np.zeros(3)
        N = lib.orc_frontend_make(npts, capi.dptr(np.ascontiguousarray(path)), capi.dptr(np.array(start)),
                                  capi.dptr(np.array(end)), capi.dptr(z3), capi.dptr(z3), fe.max_vel, fe.max_acc,
                                  fe.jps_yaw_weight, fe.jps_distance_weight, fe.trajCutLength, fe.timeResolution,
                                  fe.mintrajNum, cap, capi.dptr(inner), C.byref(initT), capi.dptr(pos), capi.dptr(ss),
                                  capi.dptr(fs), capi.dptr(fx), capi.u8ptr(cut))
        assert N == ft.TrajNum
        assert np.allclose(np.array([q[:2] for q in ft.UnOccupied_traj_pts]).reshape(-1, 2), inner[: N - 1], rtol=1e-12, atol=1e-12)
        assert math.isclose(ft.UnOccupied_initT, initT.value, rel_tol=1e-13)
        assert np.allclose(ft.start_state, ss) and np.allclose(ft.final_state, fs, rtol=1e-12)
        assert np.allclose(ft.final_state_XYTheta, fx, rtol=1e-12) and bool(cut[0]) == ft.if_cut


def test_candidate_batch_packing_roundtrip():
    geom, grid, cands = workloads.config1()
    two = capi.CandidateBatch.concat([cands, cands])
    assert two.B == 2 and two.total_pieces == 2 * cands.total_pieces
    sub = two.subset([1])
    assert np.array_equal(sub.inner_pts[: sub.total_pieces - 1], cands.inner_pts[: cands.total_pieces - 1])
    assert np.array_equal(sub.inner_init_pos, cands.inner_init_pos)
    assert two.x_offset(1) == 3 * cands.total_pieces - 1 and two.n_vars() == 2 * (3 * cands.total_pieces - 1)
