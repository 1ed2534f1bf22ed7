"""CPU tests: pin the ESDF oracle (literal restatement of sdf_map.cpp:618-715) against an
independent brute-force integer EDT + the quirk recipe of SURVEY.md Appendix B2."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from alore_legged_manipulator_b200 import capi, workloads

DBL_MAX = float(np.finfo(np.float64).max)


def brute_force_expected(occ2d: np.ndarray, gi: float, ref_compat=True):
    """Expected (dist, written-mask) for a full window = the whole (NX, NY) grid.

    Integer squared EDT by exhaustive search (no envelope), `gi*sqrt`, the reference's combine, then
    the quirk patch: local column 0, rows 1..NX-2, is the 1-D transform of the aliased column
    W(j) (j in [1, NX]): W(j) = row-pass value of cell (j, 0), W(NX) = row-pass value of (NX-1, NY-1).
    """
    NX, NY = occ2d.shape
    INF = np.int64(1) << 40

    def row_pass(seed_mask):
        g = np.full((NX, NY), INF, dtype=np.int64)
        ys = np.arange(NY)
        for x in range(NX):
            sy = ys[seed_mask[x]]
            if sy.size:
                g[x] = ((ys[:, None] - sy[None, :]) ** 2).min(axis=1)
        return g

    def col_1d(f):
        n = f.size
        xs = np.arange(n)
        c = (xs[:, None] - xs[None, :]) ** 2 + f[None, :]
        return c.min(axis=1)

    def full(g):
        return np.stack([col_1d(g[:, y]) for y in range(NY)], axis=1)

    occm = occ2d == capi.OCCUPIED
    gp, gn = row_pass(occm), row_pass(~occm)
    hp, hn = full(gp), full(gn)
    if ref_compat and NX >= 2 and NY >= 2:
        for g, h in ((gp, hp), (gn, hn)):
            f = np.concatenate([g[1:, 0], [g[NX - 1, NY - 1]]])   # f(x') = g(x'+1, 0), f(NX-1) = g(NX-1, NY-1)
            t = col_1d(f)
            h[1:, 0] = t[: NX - 1]                                   # cell (X, 0) gets the value at x = X-1

    def to_d(h):
        out = np.where(h >= INF, gi * np.sqrt(DBL_MAX), gi * np.sqrt(np.minimum(h, INF - 1).astype(np.float64)))
        return out

    pos, neg = to_d(hp), to_d(hn)
    allv = pos.copy()
    m = neg > 0.0
    allv[m] = pos[m] + (-neg[m] + gi)
    written = np.ones((NX, NY), bool)
    if ref_compat:
        written[NX - 1, :] = False
        written[:, NY - 1] = False
    return allv, written, hp, hn


SHAPES = [(48, 40), (33, 57), (64, 64), (20, 90), (70, 25), (2, 2), (3, 5), (5, 3)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dens", [0.0, 1 / 400, 1 / 23, 1 / 3, 1.0])
def test_oracle_matches_bruteforce_bit_exact(shape, dens):
    NX, NY = shape
    gi = 0.05
    rng = np.random.default_rng(NX * 1000 + NY + int(dens * 1e4))
    u = rng.random((NX, NY))
    occ = np.full((NX, NY), capi.UNOCCUPIED, np.uint8)
    occ[u < dens] = capi.OCCUPIED
    occ[(u >= dens) & (u < dens + 0.05)] = capi.UNKNOWN
    geom = workloads.make_geom(NX, NY, gi)
    dist = np.full(NX * NY, -7.0)  # sentinel for "never written"
    oracle_lib.esdf_update(geom, np.ascontiguousarray(occ.reshape(-1)), (0, 0), (NX - 1, NY - 1), dist)
    exp, written, _, _ = brute_force_expected(occ, gi)
    got = dist.reshape(NX, NY)
    assert np.array_equal(got[written], exp[written])           # bit-exact
    assert np.all(got[~written] == -7.0)                        # last row / column never written (quirk q2)


def test_oracle_sq_matches_bruteforce():
    NX, NY = 40, 36
    rng = np.random.default_rng(5)
    occ = np.where(rng.random((NX, NY)) < 0.05, capi.OCCUPIED, capi.UNOCCUPIED).astype(np.uint8)
    geom = workloads.make_geom(NX, NY, 0.1)
    dist = np.zeros(NX * NY)
    sp, sn = oracle_lib.esdf_update(geom, np.ascontiguousarray(occ.reshape(-1)), (0, 0), (NX - 1, NY - 1), dist, want_sq=True)
    _, _, hp, hn = brute_force_expected(occ, 0.1)
    S = NY - 1
    for x in range(NX):
        for y in range(S):
            assert sp[x * S + y] == (hp[x, y] if hp[x, y] < (1 << 40) else DBL_MAX)
            assert sn[x * S + y] == (hn[x, y] if hn[x, y] < (1 << 40) else DBL_MAX)


def test_oracle_window_subset_only_touches_window():
    glx, gly, gi = 60, 50, 0.1
    geom = workloads.make_geom(glx, gly, gi)
    occ = workloads.random_map(glx, gly, 3, p_occ=0.05)
    lib = oracle_lib.load()
    mn = (C.c_int * 2)()
    mx = (C.c_int * 2)()
    lib.orc_esdf_window(C.byref(geom), 0.3, -0.2, 1.55, C.cast(mn, capi.c_int32_p), C.cast(mx, capi.c_int32_p))
    mn, mx = tuple(mn), tuple(mx)
    assert 0 < mn[0] < mx[0] < glx - 1 and 0 < mn[1] < mx[1] < gly - 1
    dist = np.full(glx * gly, DBL_MAX)
    oracle_lib.esdf_update(geom, occ, mn, mx, dist)
    d = dist.reshape(glx, gly)
    inside = np.zeros((glx, gly), bool)
    inside[mn[0]:mx[0], mn[1]:mx[1]] = True   # window minus last row / col
    assert np.all(d[~inside] == DBL_MAX)
    assert np.all(d[inside] < 1e6)
    sub = occ.reshape(glx, gly)[mn[0]:mx[0] + 1, mn[1]:mx[1] + 1]
    exp, written, _, _ = brute_force_expected(sub, gi)
    assert np.array_equal(d[mn[0]:mx[0] + 1, mn[1]:mx[1] + 1][written], exp[written])


def test_bilinear_lookup_overloads():
    glx, gly, gi = 40, 40, 0.1
    geom = workloads.make_geom(glx, gly, gi)
    rng = np.random.default_rng(0)
    dist = rng.random(glx * gly) * 2.0
    lib = oracle_lib.load()
    g = np.zeros(2)
    # out of map: 1e10 (3-arg, 1-arg), 100 (2-arg), 10000 (getDistanceReal)
    far = np.array([5.0, 0.0])
    assert lib.orc_dist_grad3(C.byref(geom), capi.dptr(dist), capi.dptr(far), capi.dptr(g), 0.6) == 1e10
    assert lib.orc_dist_grad2(C.byref(geom), capi.dptr(dist), capi.dptr(far), capi.dptr(g)) == 100
    assert lib.orc_dist1(C.byref(geom), capi.dptr(dist), capi.dptr(far)) == 1e10
    assert lib.orc_dist_real(C.byref(geom), capi.dptr(dist), capi.dptr(far)) == 10000
    # cell-centre query reproduces the stored value; gradient by finite differences of the interpolant
    ix, iy = 10, 17
    c = np.array([(ix + 0.5) * gi + geom.x_lower, (iy + 0.5) * gi + geom.y_lower])
    v = lib.orc_dist_grad2(C.byref(geom), capi.dptr(dist), capi.dptr(c + 1e-9), capi.dptr(g))
    assert abs(v - dist[ix * gly + iy]) < 1e-6
    p = c + np.array([0.031, 0.047])
    v0 = lib.orc_dist_grad2(C.byref(geom), capi.dptr(dist), capi.dptr(p), capi.dptr(g))
    h = 1e-6
    g2 = np.zeros(2)
    fx = lib.orc_dist_grad2(C.byref(geom), capi.dptr(dist), capi.dptr(p + [h, 0]), capi.dptr(g2))
    fy = lib.orc_dist_grad2(C.byref(geom), capi.dptr(dist), capi.dptr(p + [0, h]), capi.dptr(g2))
    assert abs((fx - v0) / h - g[0]) < 1e-5 and abs((fy - v0) / h - g[1]) < 1e-5
    # 3-arg overload leaves grad untouched when dist > mindis
    g[:] = 123.0
    lib.orc_dist_grad3(C.byref(geom), capi.dptr(dist), capi.dptr(p), capi.dptr(g), -1.0)
    assert np.all(g == 123.0)
