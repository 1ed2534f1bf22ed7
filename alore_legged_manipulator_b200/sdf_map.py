"""Host-side mirror of the reference's `SDFmap` grid-map interface for the ESDF path.

Same method names, argument meaning and error behaviour as
planning_ddr_opt/utils/plan_env/include/plan_env/sdf_map.h:96-262 (sdf_map.cpp line numbers below);
`updateESDF2d()` forwards to the C ABI (`alore_esdf_update`), everything else is the unchanged
host logic the reference keeps on the CPU (index math, painting, nearest-cell reads of the host mirror).
The C++ twin of this file is csrc/host/sdf_map.hpp.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref

import numpy as np

from . import capi

DBL_MAX = float(np.finfo(np.float64).max)


class SDFmap:
    Unknown, Unoccupied, Occupied = 0, 1, 2  # sdf_map.h:98

    def __init__(self, ctx: capi.Context, *, gridmap_interval=0.1, detection_range=5.0, global_x_lower=-10.0,
                 global_x_upper=10.0, global_y_lower=-10.0, global_y_upper=10.0, ref_compat=True, pin_host=True):
        # sdf_map.h:120-160
        self.ctx = ctx
        self.grid_interval_ = float(gridmap_interval)
        self.inv_grid_interval_ = 1 / self.grid_interval_
        self.detection_range_ = float(detection_range)
        self.global_x_lower_, self.global_x_upper_ = float(global_x_lower), float(global_x_upper)
        self.global_y_lower_, self.global_y_upper_ = float(global_y_lower), float(global_y_upper)
        self.GLX_SIZE_ = int(math.ceil((self.global_x_upper_ - self.global_x_lower_) / self.grid_interval_))
        self.GLY_SIZE_ = int(math.ceil((self.global_y_upper_ - self.global_y_lower_) / self.grid_interval_))
        self.GLXY_SIZE_ = self.GLX_SIZE_ * self.GLY_SIZE_
        self.gridmap_ = np.full(self.GLXY_SIZE_, self.Unknown, dtype=np.uint8)
        self.distance_buffer_all_ = np.full(self.GLXY_SIZE_, DBL_MAX, dtype=np.float64)
        self.odom_pos_ = np.zeros(3)
        self.has_map_ = False
        self.has_esdf_ = False
        self.esdf_need_update_ = False
        self.ref_compat = bool(ref_compat)
        # The map owns gridmap_ / distance_buffer_all_ for its whole life (like the reference's SDFmap),
        # so it may page-lock them; released in close() / at garbage collection, before numpy frees them.
        g = self.geom()
        ctx.check(ctx.lib.alore_esdf_reset(ctx.h, C.byref(g)))
        ctx.map_owner = weakref.ref(self)      # one alore_ctx holds ONE device-resident grid map
        self._pinned = []
        if pin_host:
            for a in (self.gridmap_, self.distance_buffer_all_):
                if ctx.lib.alore_host_register(ctx.h, a.ctypes.data, a.nbytes) == 0:
                    self._pinned.append(a)
        self._fin = weakref.finalize(self, SDFmap._release, ctx, list(self._pinned))

    @staticmethod
    def _release(ctx, arrays):
        for a in arrays:
            if ctx.h:
                ctx.lib.alore_host_unregister(ctx.h, a.ctypes.data)

    def close(self):
        self._fin()

    # ---- geometry ----------------------------------------------------------------------
    def geom(self) -> capi.MapGeom:
        return capi.MapGeom(self.GLX_SIZE_, self.GLY_SIZE_, self.global_x_lower_, self.global_y_lower_,
                            self.global_x_upper_, self.global_y_upper_, self.grid_interval_, self.inv_grid_interval_)

    def gridIndex2coordd(self, x, y):  # sdf_map.cpp:453-465
        return (((float(x) + 0.5) * self.grid_interval_ + self.global_x_lower_),
                ((float(y) + 0.5) * self.grid_interval_ + self.global_y_lower_))

    def coord2gridIndex(self, pt):  # sdf_map.cpp:467-472
        ix = min(max(int((pt[0] - self.global_x_lower_) * self.inv_grid_interval_), 0), self.GLX_SIZE_ - 1)
        iy = min(max(int((pt[1] - self.global_y_lower_) * self.inv_grid_interval_), 0), self.GLY_SIZE_ - 1)
        return ix, iy

    def Index2Vectornum(self, x, y):  # sdf_map.cpp:525-527
        return x * self.GLY_SIZE_ + y

    # ---- painting (float-cast quirk of sdf_map.cpp:485-509 kept) ----------------------------
    def _paint(self, coord, state):
        cx, cy = float(np.float32(coord[0])), float(np.float32(coord[1]))
        if cx < self.global_x_lower_ or cy < self.global_y_lower_ or cx >= self.global_x_upper_ or cy >= self.global_y_upper_:
            return
        ix = int((cx - self.global_x_lower_) * self.inv_grid_interval_)
        iy = int((cy - self.global_y_lower_) * self.inv_grid_interval_)
        self.gridmap_[ix * self.GLY_SIZE_ + iy] = state
        self.has_map_ = True
        self.esdf_need_update_ = True

    def setObs(self, coord):
        self._paint(coord, self.Occupied)

    def setFree(self, coord):
        self._paint(coord, self.Unoccupied)

    def isOccupied(self, ix, iy):
        return self.gridmap_[self.Index2Vectornum(ix, iy)] == self.Occupied

    # ---- ESDF -----------------------------------------------------------------------------
    def _check_owner(self):
        owner = getattr(self.ctx, "map_owner", None)
        if owner is None or owner() is not self:
            raise capi.AloreError("this Context's device map now belongs to another SDFmap (one context = one map)")

    def esdf_window(self):
        """min_esdf / max_esdf exactly as sdf_map.cpp:619-621 computes them (FP, then truncation)."""
        ox, oy = float(self.odom_pos_[0]), float(self.odom_pos_[1])
        r, inv = self.detection_range_, self.inv_grid_interval_
        mn = (int(math.floor(max(0.0, ox - r - self.global_x_lower_) * inv)),
              int(math.floor(max(0.0, oy - r - self.global_y_lower_) * inv)))
        mx = (int(math.ceil(min(self.global_x_upper_ - self.global_x_lower_, ox + r - self.global_x_lower_) * inv) - 1),
              int(math.ceil(min(self.global_y_upper_ - self.global_y_lower_, oy + r - self.global_y_lower_) * inv) - 1))
        return mn, mx

    def updateESDF2d(self):  # sdf_map.cpp:618-680 -> alore_esdf_update
        self._check_owner()
        mn, mx = self.esdf_window()
        g = self.geom()
        rc = self.ctx.lib.alore_esdf_update(self.ctx.h, C.byref(g), capi.u8ptr(self.gridmap_), mn[0], mn[1], mx[0],
                                            mx[1], capi.dptr(self.distance_buffer_all_), 1 if self.ref_compat else 0)
        self.ctx.check(rc)

    def forceUpdateESDF(self):  # sdf_map.cpp:511-516
        if not self.has_map_:
            return
        self.esdf_need_update_ = True
        self.updateESDF2d()
        self.has_esdf_ = True

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self.ctx.check(self.ctx.lib.alore_esdf_last_kernel_ms(self.ctx.h, C.byref(ms)))
        return float(ms.value)

    def last_squared(self):
        """Integer squared distances (pos, neg) of the last update's window, window-local [NX, NY]."""
        mn, mx = self.esdf_window()
        nx, ny = mx[0] - mn[0] + 1, mx[1] - mn[1] + 1
        pos = np.zeros(nx * ny, np.int32)
        neg = np.zeros(nx * ny, np.int32)
        self.ctx.check(self.ctx.lib.alore_esdf_last_sq(self.ctx.h, capi.iptr(pos), capi.iptr(neg)))
        return pos.reshape(nx, ny), neg.reshape(nx, ny)

    # ---- host reads of the mirror (JPS / minco_plan use these on the CPU) -------------------
    def _out(self, pos):
        return (pos[0] < self.global_x_lower_ or pos[1] < self.global_y_lower_ or pos[0] > self.global_x_upper_
                or pos[1] > self.global_y_upper_)

    def getDistanceReal(self, pos):  # sdf_map.cpp:865-871
        if self._out(pos):
            return 10000.0
        ix, iy = self.coord2gridIndex(pos)
        return float(self.distance_buffer_all_[ix * self.GLY_SIZE_ + iy])

    def isOccWithSafeDis(self, ix, iy, safe_dis):  # sdf_map.cpp:942-948
        return bool(self.distance_buffer_all_[self.Index2Vectornum(ix, iy)] < safe_dis)

    def getDistWithGradBilinear(self, pos, mindis=None):
        """(dist, grad) — sdf_map.cpp:760-834.  With `mindis` (3-argument overload) the out-of-map
        value is 1e10 and grad is returned as None when dist > mindis (the reference leaves it
        untouched); without it (2-argument overload) the out-of-map value is 100."""
        far = 1e10 if mindis is not None else 100.0
        if self._out(pos):
            return far, (0.0, 0.0)
        inv = self.inv_grid_interval_
        ix = min(max(int((pos[0] - self.global_x_lower_) * inv - 0.5), 0), self.GLX_SIZE_ - 1)
        iy = min(max(int((pos[1] - self.global_y_lower_) * inv - 0.5), 0), self.GLY_SIZE_ - 1)
        if ix >= self.GLX_SIZE_ - 1 or iy >= self.GLY_SIZE_ - 1:
            return far, (0.0, 0.0)
        cx, cy = self.gridIndex2coordd(ix, iy)
        dx, dy = (pos[0] - cx) * inv, (pos[1] - cy) * inv
        d = self.distance_buffer_all_
        G = self.GLY_SIZE_
        v00, v01 = float(d[ix * G + iy]), float(d[ix * G + iy + 1])
        v10, v11 = float(d[(ix + 1) * G + iy]), float(d[(ix + 1) * G + iy + 1])
        v0 = (1 - dx) * v00 + dx * v10
        v1 = (1 - dx) * v01 + dx * v11
        dist = (1 - dy) * v0 + dy * v1
        if mindis is not None and dist > mindis:
            return dist, None
        gy = (v1 - v0) * inv
        gx = ((1 - dy) * (v10 - v00) + dy * (v11 - v01)) * inv
        return dist, (gx, gy)
