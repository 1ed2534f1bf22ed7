"""Synthetic, seeded workloads for the BASELINE.json configs.

Everything is generated with numpy's PCG64 (`np.random.default_rng(seed)`) so that the CUDA path,
the oracle and the benchmarks see bit-identical inputs; nothing here reads /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

from . import front_end
from .capi import OCCUPIED, UNKNOWN, UNOCCUPIED, CandidateBatch, MapGeom


def make_geom(glx: int, gly: int, gi: float, x_lower: float | None = None, y_lower: float | None = None) -> MapGeom:
    xl = -0.5 * glx * gi if x_lower is None else x_lower
    yl = -0.5 * gly * gi if y_lower is None else y_lower
    return MapGeom(glx, gly, xl, yl, xl + glx * gi, yl + gly * gi, gi, 1 / gi)


def random_map(glx, gly, seed, p_occ=0.02, p_unknown=0.01, wall=True, boxes=0, box_cells=(6, 20)) -> np.ndarray:
    """uint8 grid, index x*gly + y.  Bernoulli(p_occ) Occupied, p_unknown Unknown, optional 1-cell
    boundary wall and `boxes` random axis-aligned boxes."""
    rng = np.random.default_rng(seed)
    u = rng.random((glx, gly))
    g = np.full((glx, gly), UNOCCUPIED, dtype=np.uint8)
    g[u < p_occ] = OCCUPIED
    g[(u >= p_occ) & (u < p_occ + p_unknown)] = UNKNOWN
    for _ in range(boxes):
        w, h = rng.integers(box_cells[0], box_cells[1] + 1, size=2)
        x0 = int(rng.integers(1, max(2, glx - w - 1)))
        y0 = int(rng.integers(1, max(2, gly - h - 1)))
        g[x0:x0 + w, y0:y0 + h] = OCCUPIED
    if wall:
        g[0, :] = g[-1, :] = OCCUPIED
        g[:, 0] = g[:, -1] = OCCUPIED
    return np.ascontiguousarray(g.reshape(-1))


def clear_disc(grid: np.ndarray, geom: MapGeom, xy, radius_m: float):
    """Marks the cells within radius of xy Unoccupied (keeps starts/goals collision-free)."""
    g = grid.reshape(geom.glx, geom.gly)
    cx = (xy[0] - geom.x_lower) * geom.inv_grid_interval
    cy = (xy[1] - geom.y_lower) * geom.inv_grid_interval
    r = radius_m * geom.inv_grid_interval
    x0, x1 = max(1, int(cx - r)), min(geom.glx - 1, int(cx + r) + 1)
    y0, y1 = max(1, int(cy - r)), min(geom.gly - 1, int(cy + r) + 1)
    xs, ys = np.meshgrid(np.arange(x0, x1), np.arange(y0, y1), indexing="ij")
    m = (xs + 0.5 - cx) ** 2 + (ys + 0.5 - cy) ** 2 <= r * r
    g[x0:x1, y0:y1][m] = UNOCCUPIED


def config1(seed=1):
    """200x200 @0.05 m map (walls + 12 boxes 0.3-1.0 m), start (-4,-4,0) -> goal (4,4,pi/2), one dog-leg path."""
    glx = gly = 200
    gi = 0.05
    geom = make_geom(glx, gly, gi)
    grid = random_map(glx, gly, seed, p_occ=0.0, p_unknown=0.0, wall=True, boxes=12, box_cells=(6, 20))
    start, goal = (-4.0, -4.0, 0.0), (4.0, 4.0, math.pi / 2)
    for p in (start, goal, (-4.0, 4.0), (0.0, 0.0)):
        clear_disc(grid, geom, p, 0.9)
    path = [(start[0], start[1]), (0.5, -0.5), (goal[0], goal[1])]
    ft = front_end.make_flat_traj(path, start, goal)
    return geom, grid, front_end.pack_candidates([ft])


def corridor_map(glx, gly, seed, width_cells=64, clutter=0.02):
    """Free corridor of `width_cells` along x through the middle, Occupied elsewhere, sparse clutter inside."""
    rng = np.random.default_rng(seed)
    g = np.full((glx, gly), OCCUPIED, dtype=np.uint8)
    y0 = gly // 2 - width_cells // 2
    g[1:-1, y0:y0 + width_cells] = UNOCCUPIED
    u = rng.random((glx, width_cells))
    blk = g[:, y0:y0 + width_cells]
    blk[(u < clutter) & (blk == UNOCCUPIED)] = OCCUPIED
    return np.ascontiguousarray(g.reshape(-1))


def free_points(grid, geom: MapGeom, dist: np.ndarray, n, seed, min_clear=1.0, margin_m=1.5):
    """n random (x, y) in cells whose ESDF value is >= min_clear, at least margin_m from the border."""
    rng = np.random.default_rng(seed)
    d = dist.reshape(geom.glx, geom.gly)
    m = int(margin_m * geom.inv_grid_interval)
    ok = np.argwhere((d[m:geom.glx - m, m:geom.gly - m] >= min_clear) & (d[m:geom.glx - m, m:geom.gly - m] < 1e9))
    if len(ok) < n:
        raise ValueError("map too cluttered for the requested clearance")
    sel = ok[rng.choice(len(ok), size=n, replace=False)] + m
    return np.stack([(sel[:, 0] + 0.5) * geom.grid_interval + geom.x_lower,
                     (sel[:, 1] + 0.5) * geom.grid_interval + geom.y_lower], axis=1)


def leg_candidates(points_xy: np.ndarray, headings=(0.0, math.pi / 2, math.pi, 3 * math.pi / 2), start_heading=0.0,
                   fe: front_end.FrontEndParams | None = None, max_legs: int | None = None, dogleg=0.0, seed=0):
    """All ordered legs i->j (i != j) x goal headings as FlatTrajData built by the front-end time
    allocation from straight (or slightly bent, `dogleg` m) paths — BASELINE config 4."""
    rng = np.random.default_rng(seed)
    fts = []
    n = len(points_xy)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            a, b = points_xy[i], points_xy[j]
            for h in headings:
                path = [tuple(a), tuple(b)]
                if dogleg > 0:
                    mid = 0.5 * (a + b)
                    d = b - a
                    nrm = np.array([-d[1], d[0]]) / (np.linalg.norm(d) + 1e-12)
                    path = [tuple(a), tuple(mid + nrm * dogleg * (2 * rng.random() - 1)), tuple(b)]
                fts.append(front_end.make_flat_traj(path, (a[0], a[1], start_heading), (b[0], b[1], h), fe))
                if max_legs is not None and len(fts) >= max_legs:
                    return front_end.pack_candidates(fts)
    return front_end.pack_candidates(fts)


def minco_coeffs(head, tail, inPs, T):
    """Dense float64 solve of the MINCO_S3NU system (gcopter/minco.hpp:817-898) — INPUT GENERATOR for the
    coefficient-space penalty workload (config 3), not part of the product path."""
    from scipy.linalg import solve_banded
    N = len(T)
    n = 6 * N
    ab = np.zeros((13, n))

    def put(i, j, v):
        ab[6 + i - j, j] = v
    b = np.zeros((n, 2))
    put(0, 0, 1.0), put(1, 1, 1.0), put(2, 2, 2.0)
    b[0:3] = np.asarray(head).T
    for i in range(N - 1):
        t1 = T[i]; t2 = t1 * t1; t3 = t2 * t1; t4 = t2 * t2; t5 = t4 * t1
        r = 6 * i
        put(r + 3, r + 3, 6.0), put(r + 3, r + 4, 24 * t1), put(r + 3, r + 5, 60 * t2), put(r + 3, r + 9, -6.0)
        put(r + 4, r + 4, 24.0), put(r + 4, r + 5, 120 * t1), put(r + 4, r + 10, -24.0)
        for k, v in enumerate((1.0, t1, t2, t3, t4, t5)):
            put(r + 5, r + k, v)
            put(r + 6, r + k, v)
        put(r + 6, r + 6, -1.0)
        for k, v in enumerate((1.0, 2 * t1, 3 * t2, 4 * t3, 5 * t4)):
            put(r + 7, r + 1 + k, v)
        put(r + 7, r + 7, -1.0)
        for k, v in enumerate((2.0, 6 * t1, 12 * t2, 20 * t3)):
            put(r + 8, r + 2 + k, v)
        put(r + 8, r + 8, -2.0)
        b[r + 5] = inPs[i]
    t1 = T[N - 1]; t2 = t1 * t1; t3 = t2 * t1; t4 = t2 * t2; t5 = t4 * t1
    for k, v in enumerate((1.0, t1, t2, t3, t4, t5)):
        put(n - 3, n - 6 + k, v)
    for k, v in enumerate((1.0, 2 * t1, 3 * t2, 4 * t3, 5 * t4)):
        put(n - 2, n - 5 + k, v)
    for k, v in enumerate((2.0, 6 * t1, 12 * t2, 20 * t3)):
        put(n - 1, n - 4 + k, v)
    b[n - 3:] = np.asarray(tail).T
    return solve_banded((6, 6), ab, b)


def random_spline_batch(B, N, geom: MapGeom, dist: np.ndarray, grid: np.ndarray, seed, K=8):
    """Config 3 input: B valid MINCO splines of N pieces in flat space (yaw, s), with start XY in free
    space.  Returns piece_off, coeffs [B*N,6,2], T [B*N], start_xy [B,2], final_xy [B,2]."""
    rng = np.random.default_rng(seed)
    starts = free_points(grid, geom, dist, B, seed + 1000, min_clear=0.3, margin_m=1.0)
    coeffs = np.zeros((B * N, 6, 2))
    Ts = np.zeros(B * N)
    final_xy = np.zeros((B, 2))
    for b in range(B):
        th0 = rng.uniform(-math.pi, math.pi)
        dth = rng.normal(0.0, 0.15, size=N)
        ds = rng.uniform(0.2, 0.6, size=N)
        th = th0 + np.cumsum(dth)
        s = np.cumsum(ds)
        T = rng.uniform(0.3, 0.5, size=N)
        head = np.array([[th0, 0.0, 0.0], [0.0, ds[0] / T[0], 0.0]])
        tail = np.array([[th[-1], 0.0, 0.0], [s[-1], 0.0, 0.0]])
        inPs = np.stack([th[:-1], s[:-1]], axis=1)
        c = minco_coeffs(head, tail, inPs, T)
        coeffs[b * N:(b + 1) * N] = c.reshape(N, 6, 2)
        Ts[b * N:(b + 1) * N] = T
        # crude end point (only used as the ALM target): integrate with the knot headings
        xy = starts[b].copy()
        prev_th = th0
        for i in range(N):
            xy += ds[i] * np.array([math.cos(0.5 * (prev_th + th[i])), math.sin(0.5 * (prev_th + th[i]))])
            prev_th = th[i]
        final_xy[b] = xy
    piece_off = (np.arange(B + 1) * N).astype(np.int32)
    return piece_off, coeffs, Ts, np.ascontiguousarray(starts), final_xy
