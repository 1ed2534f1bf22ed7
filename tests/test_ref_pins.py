"""Pins the oracle's restatement to REFERENCE-COMPILED code, bit for bit.

oracle/_ref/libref_pins.so is built by oracle/Makefile from fragments that oracle/extract_ref.py cuts, by line
range, out of the reference where it lies: sdf_map.cpp:618-715 (updateESDF2d + fillESDF — rows E1/E2, including the
window computation q4 and the aliasing quirks q1-q3), sdf_map.cpp:453-472, :525-531, :739-871, :942-948 (the three
getDistWithGradBilinear overloads, getDistanceReal, isOccWithSafeDis and their index helpers — E3/E4),
minco.hpp:43-198 (BandedSystem — M1), optimizer.cpp:573-591 and
:1069-1106 (tau <-> T maps, backwardGradT, positiveSmoothedL1 — M5/P4).  The fragments compile against the small Eigen
stand-in in oracle/eigen_shim, which only supplies containers (no arithmetic of its own on these paths: element
access, row views evaluated element by element, Vector2i as a tuple).  CPU only; the prebuilt library travels to
the GPU box, and the tests skip when neither the library nor the reference tree is present.
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from alore_legged_manipulator_b200 import capi, workloads

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libref_pins.so"
dp, u8p = capi.c_double_p, capi.c_uint8_p


@pytest.fixture(scope="module")
def ref():
    oracle_lib.load()            # builds oracle/ (and oracle/_ref when /root/reference exists)
    if not LIB.exists():
        pytest.skip("oracle/_ref/libref_pins.so not built (no reference tree here)")
    lib = C.CDLL(str(LIB))
    G = C.POINTER(capi.MapGeom)
    lib.ref_esdf_update.argtypes = [G, u8p, C.c_double, C.c_double, C.c_double, dp]
    lib.ref_banded.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp]
    lib.ref_tmaps.argtypes = [C.c_int, dp, C.c_int, dp, dp]
    lib.ref_smoothed_l1.argtypes = [C.c_double, C.c_double, dp, dp]
    lib.ref_dist_lookups.argtypes = [G, dp, C.c_int, dp, C.c_int, C.c_double, dp, dp]
    return lib


@pytest.fixture(scope="module")
def orc():
    lib = oracle_lib.load()
    lib.orc_banded.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp]
    lib.orc_tmaps.argtypes = [C.c_int, dp, C.c_int, dp, dp]
    lib.orc_smoothed_l1.argtypes = [C.c_double, C.c_double, dp, dp]
    return lib


SHAPES = [(48, 40), (33, 57), (64, 64), (20, 90), (70, 25), (40, 40), (17, 23), (96, 31)]
DENS = [1 / 3, 1 / 23, 1 / 100, 1 / 400, 0.0]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dens", DENS)
def test_esdf_oracle_equals_reference_compiled(ref, orc, shape, dens):
    """40 shapes x densities, all three cell states, full and partial windows, stale cells outside the window."""
    glx, gly = shape
    seed = int(1000 * dens) + glx * 7 + gly
    grid = workloads.random_map(glx, gly, seed, p_occ=dens, p_unknown=0.05 if dens else 0.0, wall=bool(seed & 1))
    # 1/16 m cells: bounds and (upper - lower) * inv are exact in binary, so the reference's window (ceil(..) - 1) stays
    # inside the grid (with 0.05 m, rounding can push it one row past the array the reference indexes)
    gi = 0.0625
    geom = workloads.make_geom(glx, gly, gi)
    rng = np.random.default_rng(seed)
    for odom, rng_m in (((0.0, 0.0), 1e6), ((0.3 * glx * gi - 0.01, -0.2 * gly * gi), 0.31 * min(glx, gly) * gi)):
        stale = rng.uniform(-3.0, 3.0, glx * gly)
        a, b = stale.copy(), stale.copy()
        ref.ref_esdf_update(C.byref(geom), capi.u8ptr(grid), odom[0], odom[1], rng_m, capi.dptr(a))
        mn, mx = (C.c_int * 2)(), (C.c_int * 2)()
        orc.orc_esdf_window(C.byref(geom), odom[0], odom[1], rng_m, mn, mx)
        oracle_lib.esdf_update(geom, grid, (mn[0], mn[1]), (mx[0], mx[1]), b)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (shape, dens, odom)
        assert not np.array_equal(a, stale)          # the window was written


@pytest.mark.parametrize("N", [1, 2, 3, 5, 12, 40])
def test_banded_system_oracle_equals_reference_compiled(ref, orc, N):
    """Random 6N x 6N systems of bandwidth 6/6 (with structural zeros, as MINCO has), solve and solveAdj, factor data."""
    n = 6 * N
    rng = np.random.default_rng(N)
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - 6), min(n, i + 7)):
            if rng.random() < 0.7:
                A[i, j] = rng.normal()
        A[i, i] = 4.0 + rng.random()
    for mode in (0, 1):
        b0 = rng.normal(size=(n, 2))
        ba, bb = b0.copy(), b0.copy()
        fa, fb = np.zeros(13 * n), np.zeros(13 * n)
        ref.ref_banded(n, 6, 6, capi.dptr(A), capi.dptr(ba), mode, capi.dptr(fa))
        orc.orc_banded(n, 6, 6, capi.dptr(A), capi.dptr(bb), mode, capi.dptr(fb))
        assert np.array_equal(ba.view(np.uint64), bb.view(np.uint64)) and np.array_equal(fa.view(np.uint64), fb.view(np.uint64))
        x = np.linalg.solve(A if mode == 0 else A.T, b0)
        assert np.allclose(ba, x, rtol=1e-8, atol=1e-10)     # and it really solves the system


def test_time_maps_and_smoothed_l1_oracle_equals_reference_compiled(ref, orc):
    rng = np.random.default_rng(3)
    T = np.concatenate([rng.uniform(0.05, 6.0, 4000), [1.0, 1.0 + 1e-16, 0.999999999]])
    tau = np.concatenate([rng.normal(0, 2.5, 4000), [0.0, -0.0, 1e-300, -1e-300]])
    g = rng.normal(size=tau.size)
    for which, src in ((0, T), (1, tau), (2, tau)):
        a, b = np.zeros(src.size), np.zeros(src.size)
        gg = np.ascontiguousarray(g[:src.size]) if src.size <= g.size else np.resize(g, src.size)
        ref.ref_tmaps(src.size, capi.dptr(src), which, capi.dptr(gg), capi.dptr(a))
        orc.orc_tmaps(src.size, capi.dptr(src), which, capi.dptr(gg), capi.dptr(b))
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), which
    for eps in (0.01, 0.5):
        for x in np.concatenate([rng.uniform(-1e-3, 3 * eps, 2000), [0.0, eps, eps * (1 - 1e-16)]]):
            fa, da, fb, db = C.c_double(), C.c_double(), C.c_double(), C.c_double()
            ref.ref_smoothed_l1(eps, float(x), C.byref(fa), C.byref(da))
            orc.orc_smoothed_l1(eps, float(x), C.byref(fb), C.byref(db))
            assert fa.value == fb.value and da.value == db.value


def lookup_positions(geom, rng, n):
    """Inside, on cell centres / edges, on and beyond the map bounds, in the last row / column (where the reference
    refuses to interpolate)."""
    gi = geom.grid_interval
    p = np.empty((n, 2))
    p[:, 0] = rng.uniform(geom.x_lower - 0.3, geom.x_upper + 0.3, n)
    p[:, 1] = rng.uniform(geom.y_lower - 0.3, geom.y_upper + 0.3, n)
    k = n // 4
    p[:k, 0] = geom.x_lower + gi * rng.integers(0, geom.glx + 1, k)                   # cell edges
    p[k:2 * k, 1] = geom.y_lower + gi * (rng.integers(0, geom.gly, k) + 0.5)          # cell centres
    p[2 * k:2 * k + 8] = [(geom.x_lower, geom.y_lower), (geom.x_upper, geom.y_upper), (geom.x_lower, geom.y_upper),
                          (geom.x_upper - 1e-12, geom.y_lower + 1e-12), (geom.x_upper - gi, geom.y_upper - gi),
                          (geom.x_upper - 1.5 * gi, geom.y_upper - 0.5 * gi), (geom.x_lower + 0.5 * gi, geom.y_lower + 0.5 * gi),
                          (geom.x_lower - 1e-9, geom.y_lower)]
    return p


@pytest.mark.parametrize("seed", range(4))
def test_esdf_lookups_equal_reference_compiled(ref, seed):
    """E3/E4: getDistWithGradBilinear (with mindis, with gradient, value only), getDistanceReal, isOccWithSafeDis —
    oracle restatement and the host twins of SDFmap against the reference's own functions, bit for bit, including the
    gradient the 3-argument overload leaves untouched when dist > mindis."""
    rng = np.random.default_rng(40 + seed)
    glx, gly = [(64, 48), (37, 91), (120, 120), (25, 25)][seed]
    geom = workloads.make_geom(glx, gly, 0.0625)
    grid = workloads.random_map(glx, gly, seed, p_occ=0.05, p_unknown=0.02)
    dist = np.full(glx * gly, np.finfo(np.float64).max)
    oracle_lib.esdf_update(geom, grid, (0, 0), (glx - 1, gly - 1), dist)
    dist[dist > 1e300] = 7.25                      # cells the reference never writes: any finite stand-in, same for both
    pos = lookup_positions(geom, rng, 4000)
    n = len(pos)
    olib = oracle_lib.load()
    G = C.byref(geom)
    for which, mindis in ((3, 0.4), (3, 1e9), (2, 0.0), (1, 0.0), (0, 0.0), (-1, 0.35)):
        out = np.zeros(n)
        sentinel = rng.normal(size=(n, 2))
        gio = sentinel.copy()
        ref.ref_dist_lookups(G, capi.dptr(dist), n, capi.dptr(np.ascontiguousarray(pos)), which, mindis, capi.dptr(out),
                             capi.dptr(gio) if which >= 2 else None)
        for i in range(n):
            pi = np.ascontiguousarray(pos[i])
            g = sentinel[i].copy()
            if which == 3:
                v = olib.orc_dist_grad3(G, capi.dptr(dist), capi.dptr(pi), capi.dptr(g), mindis)
            elif which == 2:
                v = olib.orc_dist_grad2(G, capi.dptr(dist), capi.dptr(pi), capi.dptr(g))
            elif which == 1:
                v = olib.orc_dist1(G, capi.dptr(dist), capi.dptr(pi))
            elif which == 0:
                v = olib.orc_dist_real(G, capi.dptr(dist), capi.dptr(pi))
            else:
                ix = min(max(int((pi[0] - geom.x_lower) * geom.inv_grid_interval), 0), glx - 1)
                iy = min(max(int((pi[1] - geom.y_lower) * geom.inv_grid_interval), 0), gly - 1)
                v = 1.0 if dist[ix * gly + iy] < mindis else 0.0
            assert v == out[i], (which, i, pos[i], v, out[i])
            if which >= 2:
                assert g[0] == gio[i, 0] and g[1] == gio[i, 1], (which, i, pos[i], g, gio[i])
