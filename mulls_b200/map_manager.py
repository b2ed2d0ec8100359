"""Host-side mirror of lo::MapManager::update_local_map over the C-ABI (src/map_manager.cpp:17-145).

`LocalMap` is the device-resident counterpart of the `cblock_local_map` that test/mulls_slam.cpp:438-442 keeps
updating: its six class clouds (pc_ground ... pc_vertex) live in HBM, `MapManager.update_local_map` has the argument
list and defaults of include/pgo/map_manager.h:22-32, and `LocalMap.mm_lls_icp` is the scan-to-map registration
(block1 = the map). Everything runs on the GPU through libmulls_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .registration import CloudBlock, Context


class LocalMap:
    """A mulls_map on `ctx` (a registration.Context); at most `max_pts_per_class` points per feature class."""

    def __init__(self, ctx: Context, max_pts_per_class: int = 1 << 17):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = self.lib.mulls_map_create(ctx.handle, int(max_pts_per_class))
        if not self.handle:
            raise RuntimeError("mulls_map_create failed: " + self.lib.mulls_last_error(ctx.handle).decode())

    def close(self):
        if getattr(self, "handle", None):
            if getattr(self.ctx, "handle", None):  # the map must go before its context
                self.lib.mulls_map_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- content ------------------------------------------------------------------------------
    def set(self, clouds, pose_lo=None):
        """Replace the map by six host clouds (ground, pillar, facade, beam, roof, vertex) with pose `pose_lo`."""
        arrs = [abi.as_aos48(c) for c in clouds]
        views = (abi.CloudView * 6)(*[abi.cloud_view(a) for a in arrs])
        pose = np.ascontiguousarray(np.eye(4) if pose_lo is None else pose_lo, dtype=np.float64).reshape(16)
        self.ctx._check(self.lib.mulls_map_set(self.handle, views, pose.ctypes.data_as(C.POINTER(C.c_double))))

    def info(self) -> dict:
        info = abi.MapInfo()
        self.ctx._check(self.lib.mulls_map_get_info(self.handle, C.byref(info)))
        return abi.map_info_to_dict(info)

    def download(self):
        """The six class clouds as (n,12) float32 arrays (a D2H copy; the registration never needs it)."""
        out = []
        for c in range(6):
            n = C.c_size_t(0)
            self.ctx._check(self.lib.mulls_map_download(self.handle, c, None, 0, C.byref(n)))
            buf = np.zeros((n.value, 12), np.float32)
            if n.value:
                self.ctx._check(self.lib.mulls_map_download(self.handle, c, buf.ctypes.data_as(C.POINTER(C.c_float)),
                                                            n.value, C.byref(n)))
            out.append(buf)
        return out

    # ---- update_local_map ------------------------------------------------------------------------
    def update(self, scan_down, scan_pose_lo, params: abi.MapParams) -> dict:
        arrs = [abi.as_aos48(c) for c in scan_down]
        views = (abi.CloudView * 6)(*[abi.cloud_view(a) for a in arrs])
        pose = np.ascontiguousarray(scan_pose_lo, dtype=np.float64).reshape(16)
        info = abi.MapInfo()
        self.ctx._check(self.lib.mulls_map_update(self.handle, views, pose.ctypes.data_as(C.POINTER(C.c_double)),
                                                  C.byref(params), C.byref(info)))
        return abi.map_info_to_dict(info)

    # ---- scan-to-map registration ---------------------------------------------------------------
    def icp_run(self, src, params: abi.IcpParams, init_guess, want_trace: bool = False):
        """mm_lls_icp with block1 = this map: only the six source clouds are copied to the device."""
        arrs = [abi.as_aos48(c) for c in src]
        views = (abi.CloudView * 6)(*[abi.cloud_view(a) for a in arrs])
        init = np.ascontiguousarray(init_guess, dtype=np.float64).reshape(16)
        res = abi.IcpResult()
        tr = abi.IcpTrace() if want_trace else None
        self.ctx._check(self.lib.mulls_icp_run_to_map(self.ctx.handle, self.handle, views, C.byref(params),
                                                      init.ctypes.data_as(C.POINTER(C.c_double)), C.byref(res),
                                                      C.byref(tr) if tr is not None else None))
        return abi.result_to_dict(res), (abi.trace_to_dict(tr) if tr is not None else None)


class MapManager:
    """lo::MapManager — the local-map part (include/pgo/map_manager.h:22-32)."""

    def update_local_map(self, local_map: LocalMap, last_target_cblock: CloudBlock, last_target_pose_lo,
                         local_map_radius: float = 80, max_num_pts: int = 20000, kept_vertex_num: int = 800,
                         last_frame_reliable_radius: float = 60, map_based_dynamic_removal_on: bool = False,
                         used_feature_type: str = "111110", dynamic_removal_center_radius: float = 30.0,
                         dynamic_dist_thre_min: float = 0.3, dynamic_dist_thre_max: float = 3.0,
                         near_dist_thre: float = 0.03, recalculate_feature_on: bool = False,
                         random_seed: int = 0) -> bool:
        """Same arguments as the reference; `last_target_pose_lo` is last_target_cblock->pose_lo (CloudBlock does
        not carry poses). Unlike the reference the scan block is left untouched."""
        p = abi.default_map_params()
        p.local_map_radius = local_map_radius
        p.max_num_pts = max_num_pts
        p.kept_vertex_num = kept_vertex_num
        p.last_frame_reliable_radius = last_frame_reliable_radius
        p.map_based_dynamic_removal_on = 1 if map_based_dynamic_removal_on else 0
        p.used_feature_type = used_feature_type.encode()
        p.dynamic_removal_center_radius = dynamic_removal_center_radius
        p.dynamic_dist_thre_min = dynamic_dist_thre_min
        p.dynamic_dist_thre_max = dynamic_dist_thre_max
        p.near_dist_thre = near_dist_thre
        p.recalculate_feature_on = 1 if recalculate_feature_on else 0
        p.random_seed = random_seed
        b = last_target_cblock
        scan = (b.pc_ground_down, b.pc_pillar_down, b.pc_facade_down, b.pc_beam_down, b.pc_roof_down, b.pc_vertex)
        local_map.update(scan, last_target_pose_lo, p)
        return True
