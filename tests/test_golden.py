"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle).

The reference ships no golden vectors for this path, so the fixtures are frozen oracle outputs:
* CPU: the oracle still reproduces them bit for bit (guards the checker itself against drift);
* GPU: the CUDA path, through the C ABI, reproduces them bit for bit WITHOUT the oracle in the loop.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "golden"))
import make_golden as mg  # noqa: E402
import oracle_lib  # noqa: E402
from alore_legged_manipulator_b200 import capi  # noqa: E402

G = HERE / "golden"
DBL_MAX = np.finfo(np.float64).max


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def portable_trig():
    lib = oracle_lib.load()
    lib.orc_set_trig_portable(1)
    yield
    lib.orc_set_trig_portable(0)


# ---- CPU: the oracle against its own frozen outputs ---------------------------------------------------------------
def test_oracle_esdf_matches_golden():
    geom, grid, mn, mx = mg.esdf_case()
    dist = np.full(geom.glx * geom.gly, DBL_MAX)
    sp, sn = oracle_lib.esdf_update(geom, grid, mn, mx, dist, want_sq=True)
    g = np.load(G / "esdf_window_128x112.npz")
    assert np.array_equal(bits(dist), g["dist_bits"])
    assert np.array_equal(sp.astype(np.int64), g["pos_sq"]) and np.array_equal(sn.astype(np.int64), g["neg_sq"])


def _opt_world():
    geom, grid, cands = mg.opt_case()
    dist = np.full(geom.glx * geom.gly, DBL_MAX)
    oracle_lib.esdf_update(geom, grid, (0, 0), (geom.glx - 1, geom.gly - 1), dist)
    return geom, grid, cands, dist


def test_oracle_cost_gradient_and_plan_match_golden(portable_trig):
    geom, grid, cands, dist = _opt_world()
    prm = mg.params()
    g = np.load(G / "cost_gradient_config1.npz")
    for b in range(cands.B):
        x = oracle_lib.initial_x(cands, b)
        for stage in (0, 1):
            c, grad, err = oracle_lib.cost(prm, geom, dist, cands, b, stage, x)
            assert bits([c])[0] == g[f"cost_{b}_{stage}"][0]
            assert np.array_equal(bits(grad), g[f"grad_{b}_{stage}"]) and np.array_equal(bits(err), g[f"err_{b}_{stage}"])
    res = oracle_lib.opt_batch(prm, geom, dist, cands, 2)
    p = np.load(G / "minco_plan_config1.npz")
    assert np.array_equal(res.ok, p["ok"]) and np.array_equal(res.evals, p["evals"]) and np.array_equal(res.status, p["status"])
    assert np.array_equal(bits(res.coeffs).ravel(), p["coeffs_bits"].ravel()) and np.array_equal(bits(res.cost), p["cost_bits"])


def test_oracle_penalty_matches_golden(portable_trig):
    geom, grid, cands, dist = _opt_world()
    from alore_legged_manipulator_b200 import workloads
    po, coeffs, T, s_xy, f_xy = workloads.random_spline_batch(12, 16, geom, dist, grid, seed=3)
    c, gC, gT, err = oracle_lib.penalty_batch(mg.params(), geom, dist, po, coeffs, T, s_xy, f_xy)
    g = np.load(G / "penalty_batch_12x16.npz")
    assert np.array_equal(bits(c), g["cost_bits"]) and np.array_equal(bits(gC).ravel(), g["gC_bits"].ravel())
    assert np.array_equal(bits(gT), g["gT_bits"]) and np.array_equal(bits(err).ravel(), g["err_bits"].ravel())


# ---- GPU: the CUDA path against the same fixtures, no oracle involved ------------------------------------------------
@pytest.mark.gpu
def test_cuda_esdf_matches_golden(ctx):
    geom, grid, mn, mx = mg.esdf_case()
    ctx.check(ctx.lib.alore_esdf_reset(ctx.h, C.byref(geom)))
    dist = np.full(geom.glx * geom.gly, DBL_MAX)
    ctx.check(ctx.lib.alore_esdf_update(ctx.h, C.byref(geom), capi.u8ptr(grid), mn[0], mn[1], mx[0], mx[1], capi.dptr(dist), 1))
    g = np.load(G / "esdf_window_128x112.npz")
    assert np.array_equal(bits(dist), g["dist_bits"])


@pytest.mark.gpu
def test_cuda_optimizer_matches_golden():
    import alore_legged_manipulator_b200 as alore
    from alore_legged_manipulator_b200 import workloads
    from alore_legged_manipulator_b200.ms_planner import MSPlanner
    from test_esdf_gpu import make_sdf
    ctx = alore.Context(0)
    geom, grid, cands = mg.opt_case()
    m = make_sdf(ctx, geom.glx, geom.gly, geom.grid_interval, grid)
    m.forceUpdateESDF()
    prm = mg.params()
    pl = MSPlanner(ctx, prm, m)
    x = np.concatenate([_initial_x(cands, b) for b in range(cands.B)])
    g = np.load(G / "cost_gradient_config1.npz")
    for stage in (0, 1):
        c, grad, err = pl.cost_batch(cands, stage, x)
        off = 0
        for b in range(cands.B):
            n = 3 * int(cands.piece_off[b + 1] - cands.piece_off[b]) - 1
            assert bits([c[b]])[0] == g[f"cost_{b}_{stage}"][0]
            assert np.array_equal(bits(grad[off:off + n]), g[f"grad_{b}_{stage}"])
            if stage == 1:
                assert np.array_equal(bits(err[b]), g[f"err_{b}_{stage}"])
            off += n
    res = pl.minco_plan_batch(cands)
    p = np.load(G / "minco_plan_config1.npz")
    assert np.array_equal(res.ok, p["ok"]) and np.array_equal(res.status, p["status"]) and np.array_equal(res.alm_iters, p["alm_iters"])
    assert np.array_equal(res.evals, p["evals"] - p["replans"])      # the oracle also counts the reference's printing evaluation
    assert np.array_equal(bits(res.coeffs).ravel(), p["coeffs_bits"].ravel()) and np.array_equal(bits(res.cost), p["cost_bits"])
    assert np.array_equal(bits(res.piece_T), p["piece_T_bits"]) and np.array_equal(bits(res.tail_s), p["tail_bits"])
    po, coeffs, T, s_xy, f_xy = workloads.random_spline_batch(12, 16, m.geom(), m.distance_buffer_all_, grid, seed=3)
    c, gC, gT, err = pl.penalty_batch(po, coeffs, T, s_xy, f_xy)
    q = np.load(G / "penalty_batch_12x16.npz")
    assert np.array_equal(bits(c), q["cost_bits"]) and np.array_equal(bits(gC).ravel(), q["gC_bits"].ravel())
    assert np.array_equal(bits(gT), q["gT_bits"]) and np.array_equal(bits(err).ravel(), q["err_bits"].ravel())
    m.close()
    ctx.close()


def _initial_x(cands, b):
    """x0 = [inner points | tail s | tau] (optimizer.cpp:277-286), host-side restatement used only to feed the cost call."""
    p0, p1 = int(cands.piece_off[b]), int(cands.piece_off[b + 1])
    N = p1 - p0
    T = float(cands.init_T[b])
    tau = (np.sqrt(2.0 * T - 1.0) - 1.0) if T > 1.0 else (1.0 - np.sqrt(2.0 / T - 1.0))
    inner = np.asarray(cands.inner_pts[p0 - b:p1 - b - 1], dtype=np.float64).reshape(-1)
    return np.concatenate([inner, [float(cands.final_state[b][1][0])], np.full(N, tau)])
