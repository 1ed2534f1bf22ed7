"""Multi-GPU plumbing of the candidate batch: contiguous blocks of candidates per rank (no data-path
collective — candidates are independent given the read-only ESDF, which every rank rebuilds locally) and
ONE exchange step: an all-gather of (best cost, global candidate index) per rank over torch.distributed
(NCCL on the GPUs, gloo in the CPU tests).  SURVEY.md section 8(e)."""
from __future__ import annotations

import math

import numpy as np


def shard_range(rank: int, world: int, total: int | None = None, per_rank: int | None = None):
    """[lo, hi) of the global candidate array owned by `rank`.

    weak scaling (per_rank given): block `rank` of size per_rank;
    strong scaling (total given): `total` split into `world` contiguous, nearly equal blocks."""
    if per_rank is not None:
        return rank * per_rank, (rank + 1) * per_rank
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_blocks(piece_counts, world: int, samples_per_piece: int = 17):
    """Cost-balanced contiguous split (by sum of N*(2K+1) samples) of a candidate array: returns world+1 offsets."""
    w = np.asarray(piece_counts, dtype=np.float64) * samples_per_piece
    cum = np.concatenate([[0.0], np.cumsum(w)])
    offs = [0]
    for r in range(1, world):
        offs.append(int(np.searchsorted(cum, cum[-1] * r / world)))
    offs.append(len(w))
    return offs


def local_best(cost: np.ndarray, ok: np.ndarray, lo: int):
    """(cost, global index) of the cheapest successful candidate of this shard; (inf, -1) if none."""
    idx = np.flatnonzero(np.asarray(ok) == 1)
    if idx.size == 0:
        return math.inf, -1
    k = idx[np.argmin(np.asarray(cost)[idx])]
    return float(cost[k]), int(lo + k)


def gather_best(best_cost: float, best_idx: int, device=None):
    """All-gather (cost, index) from every rank and return the global winner (ties -> lowest index).
    Works without an initialised process group (single process)."""
    import torch
    import torch.distributed as dist
    mine = torch.tensor([best_cost, float(best_idx)], dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return best_cost, best_idx, [(best_cost, best_idx)]
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    pairs = [(float(t[0]), int(t[1])) for t in out]
    valid = [p for p in pairs if p[1] >= 0]
    if not valid:
        return math.inf, -1, pairs
    c, i = min(valid, key=lambda p: (p[0], p[1]))
    return c, i, pairs
