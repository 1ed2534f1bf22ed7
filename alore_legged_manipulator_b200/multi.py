"""Several GPUs of one box behind one object: ctypes twin of alore_multi (csrc/multi.cu).

Mirrors what the reference's single C++ caller holds (one SDFmap + one MSPlanner, plan_manager.hpp:120-123): the
occupancy grid goes to every device, candidates are cut into cost-balanced contiguous blocks, one host thread per
GPU inside the library, and the (best cost, index) pairs are all-gathered over NCCL.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class MultiPlanner:
    def __init__(self, devices):
        self.lib = capi.load_library()
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        self.h = C.c_void_p()
        rc = self.lib.alore_create_multi(arr, len(self.devices), C.byref(self.h))
        if rc != 0:
            raise capi.AloreError(f"alore_create_multi({self.devices}) failed rc={rc}")
        self.block_offsets = None

    def check(self, rc: int):
        if rc != 0:
            raise capi.AloreError(f"rc={rc}: {self.lib.alore_multi_last_error(self.h).decode()}")

    def esdf_update(self, geom: capi.MapGeom, occ: np.ndarray, mn, mx, dist: np.ndarray, ref_compat: int = 1):
        """SDFmap::updateESDF2d on every device; `dist` (the host mirror) is written once."""
        self.check(self.lib.alore_multi_esdf_update(self.h, C.byref(geom), capi.u8ptr(occ), mn[0], mn[1], mx[0], mx[1],
                                                    capi.dptr(dist), ref_compat))

    def minco_plan_batch(self, prm: capi.Params, cands: capi.CandidateBatch):
        """B x MSPlanner::minco_plan sharded over the devices -> (results, best_cost, best_idx)."""
        res = capi.ResultBatch(cands)
        cs, rs = cands.as_struct(), res.as_struct()
        bc, bi = C.c_double(), C.c_int32()
        offs = np.zeros(len(self.devices) + 1, np.int32)
        self.check(self.lib.alore_multi_opt_batch(self.h, C.byref(prm), C.byref(cs), C.byref(rs), C.byref(bc), C.byref(bi),
                                                  capi.iptr(offs)))
        self.block_offsets = offs
        return res, float(bc.value), int(bi.value)

    def close(self):
        if self.h:
            self.lib.alore_destroy_multi(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
