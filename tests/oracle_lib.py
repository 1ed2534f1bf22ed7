"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Builds the library with oracle/Makefile when it is missing.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from alore_legged_manipulator_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "liboracle.so"

dp, ip, u8p = capi.c_double_p, capi.c_int32_p, capi.c_uint8_p
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True, capture_output=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    src_m = max((ROOT / "oracle" / f).stat().st_mtime for f in ("oracle_capi.cpp", "alore_oracle.hpp"))
    if not LIB.exists() or LIB.stat().st_mtime < src_m:
        build()
    lib = C.CDLL(str(LIB))
    G, P = C.POINTER(capi.MapGeom), C.POINTER(capi.Params)
    lib.orc_esdf_window.argtypes = [G, C.c_double, C.c_double, C.c_double, ip, ip]
    lib.orc_esdf_window.restype = None
    lib.orc_esdf_update.argtypes = [G, u8p, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp]
    lib.orc_esdf_update_timed.argtypes = [G, u8p, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int]
    lib.orc_esdf_update_timed.restype = C.c_double
    for n in ("orc_dist_grad3", "orc_dist_grad2", "orc_dist1", "orc_dist_real"):
        getattr(lib, n).restype = C.c_double
    lib.orc_dist_grad3.argtypes = [G, dp, dp, dp, C.c_double]
    lib.orc_dist_grad2.argtypes = [G, dp, dp, dp]
    lib.orc_dist1.argtypes = [G, dp, dp]
    lib.orc_dist_real.argtypes = [G, dp, dp]
    lib.orc_minco_solve.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.orc_minco_adjoint.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.orc_minco_matrix.argtypes = [C.c_int, dp, dp]
    lib.orc_penalty.argtypes = [P, G, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.orc_penalty_batch.argtypes = [P, G, dp, C.c_int, ip, dp, dp, dp, dp, dp, dp, dp, dp, C.c_int]
    lib.orc_cost.argtypes = [P, G, dp, C.POINTER(capi.Candidates), C.c_int, C.c_int, dp, dp, dp, C.c_double, dp, dp, dp]
    lib.orc_initial_x.argtypes = [C.POINTER(capi.Candidates), C.c_int, dp]
    lib.orc_opt_batch.argtypes = [P, G, dp, C.POINTER(capi.Candidates), C.POINTER(capi.Results), C.c_int]
    lib.orc_final_collision.argtypes = [P, G, dp, C.c_int, dp, dp, dp, ip, dp]
    lib.orc_lbfgs_rosenbrock.argtypes = [C.c_int, dp, C.POINTER(capi.LbfgsParams), dp, ip, ip]
    lib.orc_frontend_make.argtypes = [C.c_int, dp, dp, dp, dp, dp] + [C.c_double] * 6 + [C.c_int, C.c_int, dp, dp, dp,
                                                                                       dp, dp, dp, u8p]
    lib.orc_hardware_threads.restype = C.c_int
    lib.orc_params_default.argtypes = [P]
    lib.orc_params_default.restype = None
    _lib = lib
    return lib


# ---- convenience wrappers ----------------------------------------------------------------

def default_params() -> capi.Params:
    """The reference's yaml defaults WITHOUT touching the product library (the oracle states them a second time)."""
    p = capi.Params()
    load().orc_params_default(C.byref(p))
    return p



def esdf_update(geom: capi.MapGeom, occ: np.ndarray, mn, mx, dist: np.ndarray, want_sq=False):
    lib = load()
    nxy = (mx[0] - mn[0] + 1) * (mx[1] - mn[1] + 1)
    sp = np.zeros(nxy) if want_sq else None
    sn = np.zeros(nxy) if want_sq else None
    lib.orc_esdf_update(C.byref(geom), capi.u8ptr(occ), mn[0], mn[1], mx[0], mx[1], capi.dptr(dist),
                        capi.dptr(sp) if want_sq else None, capi.dptr(sn) if want_sq else None)
    return sp, sn


def opt_batch(prm: capi.Params, geom: capi.MapGeom, dist: np.ndarray, cands: capi.CandidateBatch, nthreads=1):
    lib = load()
    res = capi.ResultBatch(cands)
    cs, rs = cands.as_struct(), res.as_struct()
    lib.orc_opt_batch(C.byref(prm), C.byref(geom), capi.dptr(dist), C.byref(cs), C.byref(rs), nthreads)
    return res


def cost(prm, geom, dist, cands: capi.CandidateBatch, b: int, stage: int, x: np.ndarray, lam=None, rho=None,
         safe_dis=None):
    lib = load()
    n = x.size
    g = np.zeros(n)
    c = C.c_double()
    err = np.zeros(2)
    cs = cands.as_struct()
    lam_a = np.ascontiguousarray(lam, dtype=np.float64) if lam is not None else None
    rho_a = np.ascontiguousarray(rho, dtype=np.float64) if rho is not None else None
    lib.orc_cost(C.byref(prm), C.byref(geom), capi.dptr(dist), C.byref(cs), b, stage, capi.dptr(np.ascontiguousarray(x)),
                 capi.dptr(lam_a) if lam_a is not None else None, capi.dptr(rho_a) if rho_a is not None else None,
                 float(prm.safeDis if safe_dis is None else safe_dis), C.byref(c), capi.dptr(g), capi.dptr(err))
    return float(c.value), g, err


def initial_x(cands: capi.CandidateBatch, b: int) -> np.ndarray:
    lib = load()
    n = 3 * int(cands.piece_off[b + 1] - cands.piece_off[b]) - 1
    x = np.zeros(n)
    cs = cands.as_struct()
    lib.orc_initial_x(C.byref(cs), b, capi.dptr(x))
    return x


def penalty_batch(prm, geom, dist, piece_off, coeffs, T, start_xy, final_xy, nthreads=1):
    lib = load()
    B = piece_off.size - 1
    tot = int(piece_off[-1])
    cost_ = np.zeros(B)
    gC = np.zeros((tot, 6, 2))
    gT = np.zeros(tot)
    err = np.zeros((B, 2))
    lib.orc_penalty_batch(C.byref(prm), C.byref(geom), capi.dptr(dist), B, capi.iptr(piece_off), capi.dptr(coeffs),
                          capi.dptr(T), capi.dptr(start_xy), capi.dptr(final_xy), capi.dptr(cost_), capi.dptr(gC),
                          capi.dptr(gT), capi.dptr(err), nthreads)
    return cost_, gC, gT, err


def minco_solve(N, head, tail, inPs, T, ew=(1.0, 1.0)):
    lib = load()
    coeffs = np.zeros((6 * N, 2))
    e = C.c_double()
    gdC = np.zeros((6 * N, 2))
    gdT = np.zeros(N)
    ins = np.ascontiguousarray(inPs, dtype=np.float64) if N > 1 else np.zeros(2)
    lib.orc_minco_solve(N, capi.dptr(np.ascontiguousarray(head, dtype=np.float64)),
                        capi.dptr(np.ascontiguousarray(tail, dtype=np.float64)), capi.dptr(ins),
                        capi.dptr(np.ascontiguousarray(T, dtype=np.float64)),
                        capi.dptr(np.asarray(ew, dtype=np.float64)), capi.dptr(coeffs), C.byref(e), capi.dptr(gdC),
                        capi.dptr(gdT))
    return coeffs, float(e.value), gdC, gdT


def final_collision(prm, geom, dist, N, coeffs, T, start_xy):
    lib = load()
    col = np.zeros(1, np.int32)
    md = C.c_double()
    lib.orc_final_collision(C.byref(prm), C.byref(geom), capi.dptr(dist), N, capi.dptr(np.ascontiguousarray(coeffs)),
                            capi.dptr(np.ascontiguousarray(T)), capi.dptr(np.ascontiguousarray(start_xy)),
                            capi.iptr(col), C.byref(md))
    return int(col[0]), float(md.value)
