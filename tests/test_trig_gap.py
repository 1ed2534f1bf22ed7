"""What `optimised trajectories within 1e-9 of the reference` can and cannot mean.

The CUDA path is bit-identical to the oracle when both use the same portable sin/cos (tests/test_optimizer_gpu.py).
The REFERENCE calls glibc's sin/cos, which differ from any other implementation by <= 1 ulp in ~7 % of the calls; a
single cost / gradient evaluation then agrees to ~1e-15 (asserted <= 1e-12 elsewhere), but the optimizer amplifies it:
L-BFGS with a bisection line search and a `past/delta` stopping rule (delta = 5e-4 relative) takes different discrete
decisions.  This file MEASURES that gap on candidates of the benchmark block and asserts bounds on its distribution:
the two runs are the same optimizer landing in the same basin — ok flags agree, final costs agree to the percent
level the stopping rule allows — not on every trajectory to 1e-9, which no implementation other than the reference
binary itself can deliver.  Measured on the 2080-candidate bench sample (profiles/trig_gap_r02.md): ok agreement
99.95 %, relative final-cost difference median 0.56 %, p90 3.3 %, p99 10 %, max 40 %, same winning candidate.
"""
import numpy as np
import pytest

import oracle_lib
from alore_legged_manipulator_b200 import capi

BOUNDS = dict(ok_agreement=0.97, median=0.02, p90=0.10, worst=1.0)


def bench_sample(n=96):
    import bench
    geom, grid = bench.build_world()
    dist = np.full(geom.glx * geom.gly, np.finfo(np.float64).max)
    oracle_lib.esdf_update(geom, grid, (0, 0), (geom.glx - 1, geom.gly - 1), dist)
    pts = bench.way_points(geom, grid, dist)
    full = bench.candidates_from_points(pts, 0, 2080)
    idx = bench.sample_indices(full.B, n, seed=5)
    return geom, grid, dist, full.subset(idx)


def gap(a, b):
    both = (a.ok == 1) & (b.ok == 1)
    rel = np.abs(a.cost - b.cost) / np.maximum(1e-300, np.abs(b.cost))
    return float((a.ok == b.ok).mean()), rel[both]


def check(a, b):
    agree, rel = gap(a, b)
    assert agree >= BOUNDS["ok_agreement"], agree
    assert np.median(rel) <= BOUNDS["median"] and np.percentile(rel, 90) <= BOUNDS["p90"] and rel.max() <= BOUNDS["worst"], \
        (np.median(rel), np.percentile(rel, 90), rel.max())
    assert rel.max() > 1e-9          # and the gap is real: identical arithmetic is the only way to 1e-9


def test_portable_vs_glibc_trig_oracle_gap_is_bounded():
    geom, grid, dist, cands = bench_sample()
    lib = oracle_lib.load()
    prm = oracle_lib.default_params()
    lib.orc_set_trig_portable(1)
    try:
        a = oracle_lib.opt_batch(prm, geom, dist, cands, 8)
    finally:
        lib.orc_set_trig_portable(0)
    b = oracle_lib.opt_batch(prm, geom, dist, cands, 8)
    check(a, b)


@pytest.mark.gpu
def test_cuda_vs_glibc_trig_oracle_gap_is_bounded(ctx):
    """The product against the oracle in its DEFAULT mode (glibc sin/cos, what the reference calls)."""
    import alore_legged_manipulator_b200 as alore
    from alore_legged_manipulator_b200.ms_planner import MSPlanner
    from test_esdf_gpu import make_sdf
    geom, grid, dist, cands = bench_sample()
    c2 = alore.Context(0)
    m = make_sdf(c2, geom.glx, geom.gly, geom.grid_interval, grid)
    m.updateESDF2d()
    prm = capi.default_params()
    res = MSPlanner(c2, prm, m).minco_plan_batch(cands)
    ref = oracle_lib.opt_batch(prm, m.geom(), m.distance_buffer_all_, cands, 8)
    check(res, ref)
    m.close()
    c2.close()
