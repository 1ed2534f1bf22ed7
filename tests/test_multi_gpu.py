"""alore_multi (csrc/multi.cu): the multi-GPU path behind the C ABI.  Rank invariance:
the same candidate set gives the identical winner and identical per-candidate bits on 1 device, through alore_multi
with 1 device, and — when the box has them — through alore_multi with 2 devices (one host thread per GPU, NCCL
all-gather of (cost, index) inside the library)."""
import numpy as np
import pytest

import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import capi, workloads
from alore_legged_manipulator_b200.ms_planner import MSPlanner
from alore_legged_manipulator_b200.multi import MultiPlanner
from test_esdf_gpu import make_sdf

pytestmark = pytest.mark.gpu


def device_count():
    import torch
    return torch.cuda.device_count()


def world_and_cands():
    glx = gly = 400
    grid = workloads.random_map(glx, gly, 7, p_occ=0.0, p_unknown=0.0, wall=True, boxes=25, box_cells=(6, 24))
    geom = workloads.make_geom(glx, gly, 0.05)
    return geom, grid


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_matches_single_device(ndev):
    if device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    geom, grid = world_and_cands()
    prm = capi.default_params()
    prm.alm_max_outer = 20
    # single device, plain API
    ctx = alore.Context(0)
    m = make_sdf(ctx, geom.glx, geom.gly, 0.05, grid)
    m.updateESDF2d()
    pts = workloads.free_points(grid, m.geom(), m.distance_buffer_all_, 9, 5, min_clear=0.8)
    cands = workloads.leg_candidates(pts, headings=(0.0, 1.57), max_legs=72)
    ref = MSPlanner(ctx, prm, m).minco_plan_batch(cands)
    dist_ref = m.distance_buffer_all_.copy()
    mn, mx = m.esdf_window()
    gm = m.geom()
    ok = np.flatnonzero(ref.ok == 1)
    want_idx = int(ok[np.argmin(ref.cost[ok])])
    m.close()
    ctx.close()
    # the same through alore_multi
    mp = MultiPlanner(list(range(ndev)))
    dist = np.full(geom.glx * geom.gly, np.finfo(np.float64).max)
    mp.esdf_update(gm, grid, mn, mx, dist)
    assert np.array_equal(dist, dist_ref)
    res, bc, bi = mp.minco_plan_batch(prm, cands)
    offs = mp.block_offsets
    assert offs[0] == 0 and offs[-1] == cands.B and np.all(np.diff(offs) > 0)
    for a, b in ((res.ok, ref.ok), (res.status, ref.status), (res.evals, ref.evals), (res.cost, ref.cost),
                 (res.coeffs, ref.coeffs), (res.piece_T, ref.piece_T), (res.inner_pts, ref.inner_pts), (res.tail_s, ref.tail_s)):
        assert np.array_equal(a, b)
    assert bi == want_idx and bc == ref.cost[want_idx]
    mp.close()
