"""Generates the golden fixtures of tests/golden/ from the CPU oracle (oracle/liboracle.so, `trig_portable` mode).

The reference ships no golden vectors, known-answer tests or fixtures for this path, and its
sources cannot be built here (ROS + Eigen + PCL), so the goldens are FROZEN ORACLE OUTPUTS: they pin the oracle against
silent drift (CPU test) and give the CUDA path a target that does not need the oracle at run time (GPU test).
Inputs are regenerated deterministically by the tests from the same seeded generators; the fixtures hold outputs only,
as raw little-endian bit patterns (uint64 views of the doubles), so equality is bit equality.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib  # noqa: E402
from alore_legged_manipulator_b200 import capi, front_end, workloads  # noqa: E402

HERE = Path(__file__).resolve().parent


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64).copy()


def esdf_case():
    """96 x 80 window inside a 128 x 112 grid (seed 21): Occupied / Unoccupied / Unknown cells, ref_compat semantics."""
    glx, gly = 128, 112
    grid = workloads.random_map(glx, gly, 21, p_occ=0.06, p_unknown=0.03, wall=True, boxes=3, box_cells=(4, 10))
    geom = workloads.make_geom(glx, gly, 0.05)
    mn, mx = (9, 16), (104, 95)
    return geom, grid, mn, mx


def opt_case():
    """BASELINE configs[0] (workloads.config1) plus three more legs on the same 200 x 200 map."""
    geom, grid, _ = workloads.config1()
    fts = [front_end.make_flat_traj([(-4.0, -4.0), (0.5, -0.5), (4.0, 4.0)], (-4.0, -4.0, 0.0), (4.0, 4.0, np.pi / 2)),
           front_end.make_flat_traj([(-4.0, 4.0), (0.0, 0.0)], (-4.0, 4.0, -0.5), (0.0, 0.0, 0.0)),
           front_end.make_flat_traj([(0.0, 0.0), (4.0, 4.0)], (0.0, 0.0, 1.0), (4.0, 4.0, 0.0)),
           front_end.make_flat_traj([(4.0, 4.0), (-4.0, -4.0)], (4.0, 4.0, 3.0), (-4.0, -4.0, 0.0))]
    return geom, grid, front_end.pack_candidates(fts)


def params():
    prm = capi.default_params()
    prm.alm_max_outer = 20
    return prm


def main():
    lib = oracle_lib.load()
    lib.orc_set_trig_portable(1)
    # ---- ESDF ---------------------------------------------------------------------------------------------------
    geom, grid, mn, mx = esdf_case()
    dist = np.full(geom.glx * geom.gly, np.finfo(np.float64).max)
    sp, sn = oracle_lib.esdf_update(geom, grid, mn, mx, dist, want_sq=True)
    np.savez_compressed(HERE / "esdf_window_128x112.npz", dist_bits=bits(dist), pos_sq=sp.astype(np.int64), neg_sq=sn.astype(np.int64))
    # ---- cost / gradient at the initial point, both stages ----------------------------------------------------------
    geom, grid, cands = opt_case()
    dist = np.full(geom.glx * geom.gly, np.finfo(np.float64).max)
    oracle_lib.esdf_update(geom, grid, (0, 0), (geom.glx - 1, geom.gly - 1), dist)
    prm = params()
    out = {}
    for b in range(cands.B):
        x = oracle_lib.initial_x(cands, b)
        for stage in (0, 1):
            c, g, err = oracle_lib.cost(prm, geom, dist, cands, b, stage, x)
            out[f"cost_{b}_{stage}"] = bits([c])
            out[f"grad_{b}_{stage}"] = bits(g)
            out[f"err_{b}_{stage}"] = bits(err)
    np.savez_compressed(HERE / "cost_gradient_config1.npz", **out)
    # ---- full minco_plan ----------------------------------------------------------------------------------------
    res = oracle_lib.opt_batch(prm, geom, dist, cands, 1)
    np.savez_compressed(HERE / "minco_plan_config1.npz", ok=res.ok, status=res.status, replans=res.replans, alm_iters=res.alm_iters,
                        evals=res.evals, cost_bits=bits(res.cost), coeffs_bits=bits(res.coeffs), piece_T_bits=bits(res.piece_T),
                        inner_bits=bits(res.inner_pts), tail_bits=bits(res.tail_s))
    # ---- coefficient-space penalty (configs[2] shape, small) -----------------------------------------------------------
    po, coeffs, T, s_xy, f_xy = workloads.random_spline_batch(12, 16, geom, dist, grid, seed=3)
    c, gC, gT, err = oracle_lib.penalty_batch(prm, geom, dist, po, coeffs, T, s_xy, f_xy)
    np.savez_compressed(HERE / "penalty_batch_12x16.npz", cost_bits=bits(c), gC_bits=bits(gC), gT_bits=bits(gT), err_bits=bits(err))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
