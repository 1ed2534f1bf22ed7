"""Front-end time allocation: way-points -> FlatTrajData (the PRODUCER of the hot path's input).

Host-side restatement of JPSPlanner::getSampleTraj / getTrajsWithTime / evaluateDuration /
evaluateLength (planning_ddr_opt/front_end/src/jps_planner/jps_planner.cpp:217-441); SURVEY.md §8f
rank 1 ("next" row).  Pure Python, used to build candidate batches from straight / dog-leg paths;
the graph search itself (JPS) stays on the reference's CPU side.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .capi import CandidateBatch


@dataclass
class FrontEndParams:
    max_vel: float = 3.0            # plan_tester/config/car3ms.yaml
    max_acc: float = 2.0
    jps_yaw_weight: float = 0.30    # front_end/config/jps3ms.yaml
    jps_distance_weight: float = 1.40
    trajCutLength: float = 600.0    # back_end/config/global_planning3ms.yaml
    timeResolution: float = 0.4
    mintrajNum: int = 3


@dataclass
class FlatTrajData:  # front_end/include/front_end/traj_representation.h:46-76
    UnOccupied_traj_pts: list = field(default_factory=list)    # (yaw, s, t)
    UnOccupied_initT: float = 0.0
    UnOccupied_positions: list = field(default_factory=list)   # (x, y, yaw)
    start_state: np.ndarray = field(default_factory=lambda: np.zeros((2, 3)))
    final_state: np.ndarray = field(default_factory=lambda: np.zeros((2, 3)))
    start_state_XYTheta: np.ndarray = field(default_factory=lambda: np.zeros(3))
    final_state_XYTheta: np.ndarray = field(default_factory=lambda: np.zeros(3))
    if_cut: bool = False

    @property
    def TrajNum(self) -> int:
        return len(self.UnOccupied_traj_pts) + 1


def _normalize_angle(ref, ang):  # jps_planner.cpp:368-375
    while ref - ang > math.pi:
        ang += 2 * math.pi
    while ref - ang < -math.pi:
        ang -= 2 * math.pi
    return ang


def evaluate_duration(p: FrontEndParams, length, startV, endV, maxV, maxA):  # jps_planner.cpp:378-399
    sv2, ev2, mv2 = startV ** 2, endV ** 2, maxV ** 2
    if startV > maxV:
        sv2 = mv2
    if endV > p.max_vel:
        ev2 = mv2
    crit = (mv2 - sv2) / (2 * maxA) + (mv2 - ev2) / (2 * maxA)
    if length >= crit:
        return (maxV - startV) / maxA + (maxV - endV) / maxA + (length - crit) / maxV
    tmpv = math.sqrt(0.5 * (sv2 + ev2 + 2 * maxA * length))
    return (tmpv - startV) / maxA + (tmpv - endV) / maxA


def evaluate_length(p: FrontEndParams, curt, locallength, startV, endV, maxV, maxA):  # jps_planner.cpp:404-441
    sv2, ev2, mv2 = startV ** 2, endV ** 2, maxV ** 2
    if startV > maxV:
        sv2 = mv2
    if endV > p.max_vel:
        ev2 = mv2
    crit = (mv2 - sv2) / (2 * maxA) + (mv2 - ev2) / (2 * maxA)
    if locallength >= crit:
        t1 = (maxV - startV) / maxA
        t2 = t1 + (locallength - crit) / maxV
        if curt <= t1:
            return startV * curt + 0.5 * maxA * curt ** 2
        if curt <= t2:
            return startV * t1 + 0.5 * maxA * t1 ** 2 + (curt - t1) * maxV
        return startV * t1 + 0.5 * maxA * t1 ** 2 + (t2 - t1) * maxV + maxV * (curt - t2) - 0.5 * maxA * (curt - t2) ** 2
    tmpv = math.sqrt(0.5 * (sv2 + ev2 + 2 * maxA * locallength))
    tmpt = (tmpv - startV) / maxA
    if curt <= tmpt:
        return startV * curt + 0.5 * maxA * curt ** 2
    return startV * tmpt + 0.5 * maxA * tmpt ** 2 + tmpv * (curt - tmpt) - 0.5 * maxA * (curt - tmpt) ** 2


def make_flat_traj(path, start_state, end_state, p: FrontEndParams | None = None, VAJ=(0.0, 0.0, 0.0),
                   OAJ=(0.0, 0.0, 0.0)) -> FlatTrajData:
    """path: >= 2 (x, y) way-points; start/end_state: (x, y, yaw)."""
    p = p or FrontEndParams()
    sx, sy, sth = (float(v) for v in start_state)
    samp = [[sx, sy, sth, 0.0, 0.0]]                                  # getSampleTraj, jps_planner.cpp:217-262
    th = math.atan2(path[1][1] - path[0][1], path[1][0] - path[0][0])
    th = _normalize_angle(sth, th)
    samp.append([sx, sy, th, th - sth, 0.0])
    th = math.atan2(path[0][1] - path[1][1], path[0][0] - path[1][0]) + math.pi
    th = _normalize_angle(sth, th)
    samp.append([sx, sy, th, th - sth, 0.0])
    n = len(path)
    for i in range(1, n - 1):
        pt = path[i]
        bk = samp[-1]
        samp.append([pt[0], pt[1], bk[2], 0.0, math.sqrt((pt[0] - bk[0]) ** 2 + (pt[1] - bk[1]) ** 2)])
        th = math.atan2(path[i + 1][1] - path[i][1], path[i + 1][0] - path[i][0])
        th = _normalize_angle(samp[-1][2], th)
        samp.append([pt[0], pt[1], th, th - samp[-1][2], 0.0])
    pt = path[-1]
    bk = samp[-1]
    samp.append([pt[0], pt[1], bk[2], 0.0, math.sqrt((pt[0] - bk[0]) ** 2 + (pt[1] - bk[1]) ** 2)])
    th = _normalize_angle(samp[-1][2], float(end_state[2]))
    samp.append([pt[0], pt[1], th, th - samp[-1][2], 0.0])

    cut = [samp[0]]                                                   # getTrajsWithTime, jps_planner.cpp:264-366
    plen, wlen = [0.0], [0.0]
    allw = alllen = 0.0
    if_cut = False
    cut_state = list(samp[-1][:3])
    for idx in range(1, len(samp)):
        pn = samp[idx]
        if alllen + abs(pn[4]) >= p.trajCutLength and pn[4] != 0:
            if_cut = True
            fs = samp[idx - 1]
            cut_state = [fs[t] + (pn[t] - fs[t]) * (p.trajCutLength - alllen) / abs(pn[4]) for t in range(3)]
            s5 = [cut_state[0], cut_state[1], cut_state[2], (p.trajCutLength - alllen) / abs(pn[4]) * pn[3],
                  p.trajCutLength - alllen]
            cut.append(s5)
            alllen += s5[4]
            plen.append(alllen)
            allw += p.jps_yaw_weight * abs(s5[3]) + p.jps_distance_weight * abs(s5[4])
            wlen.append(allw)
            break
        cut.append(pn)
        alllen += pn[4]
        plen.append(alllen)
        allw += p.jps_yaw_weight * abs(pn[3]) + p.jps_distance_weight * abs(pn[4])
        wlen.append(allw)
    total_t = evaluate_duration(p, allw, VAJ[0], 0.0, p.max_vel, p.max_acc)
    ft = FlatTrajData()
    sampletime = total_t / max(int(total_t / p.timeResolution + 0.5), p.mintrajNum)
    node = 1
    nn = len(cut)
    samplet = sampletime
    while samplet < total_t - 1e-3:
        arc = evaluate_length(p, samplet, allw, VAJ[0], 0.0, p.max_vel, p.max_acc)
        for k in range(node, nn):
            pn, pp = cut[k], cut[k - 1]
            tmparc = wlen[k]
            if tmparc >= arc:
                node = k
                l1 = tmparc - arc
                l = wlen[k] - wlen[k - 1]
                s_i = plen[k - 1] + (l - l1) / l * pn[4]
                yaw_i = cut[k - 1][2] + (l - l1) / l * pn[3]
                ft.UnOccupied_traj_pts.append((yaw_i, s_i, samplet))
                ft.UnOccupied_positions.append((l1 / l * pp[0] + (l - l1) / l * pn[0],
                                                l1 / l * pp[1] + (l - l1) / l * pn[1], yaw_i))
                break
        samplet += sampletime
    ft.start_state = np.array([[cut[0][2], OAJ[0], OAJ[1]], [0.0, VAJ[0], VAJ[1]]])
    ft.final_state = np.array([[cut[nn - 1][2], 0.0, 0.0], [plen[nn - 1], 0.0, 0.0]])
    ft.UnOccupied_initT = sampletime
    ft.start_state_XYTheta = np.array([sx, sy, sth])
    ft.final_state_XYTheta = np.array(cut_state, dtype=np.float64)
    ft.if_cut = if_cut
    return ft


def pack_candidates(fts) -> CandidateBatch:
    """list[FlatTrajData] -> structure-of-arrays batch (include/alore_b200.h: alore_candidates_t)."""
    po = [0]
    inner, pos = [], []
    for ft in fts:
        po.append(po[-1] + ft.TrajNum)
        for q in ft.UnOccupied_traj_pts:
            inner.append((q[0], q[1]))
        for q in ft.UnOccupied_positions:
            pos.append(tuple(q))
        pos.append(tuple(ft.final_state_XYTheta))   # optimizer.cpp:234-235
    return CandidateBatch(
        po, np.array(inner, dtype=np.float64).reshape(-1, 2), [ft.UnOccupied_initT for ft in fts],
        np.array(pos, dtype=np.float64).reshape(-1, 3), np.stack([ft.start_state for ft in fts]),
        np.stack([ft.final_state for ft in fts]), np.stack([ft.start_state_XYTheta for ft in fts]),
        np.stack([ft.final_state_XYTheta for ft in fts]), [1 if ft.if_cut else 0 for ft in fts])
