"""ctypes wrapper of oracle/libmulls_oracle.so (TEST INFRASTRUCTURE — never imported by mulls_b200/).

PARITY UNPINNED: see the header of mulls_oracle.cpp. The oracle re-uses the POD structs of the
public ABI (include/mulls_b200/abi.h via mulls_b200.abi) so that inputs/outputs are interchangeable
with the CUDA path's.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mulls_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmulls_oracle.so")
_LIB = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mulls_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "mulls_b200", "abi.h")
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in (src, hdr) if os.path.exists(f))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def load() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.orc_icp_run.restype = C.c_int
        lib.orc_icp_run.argtypes = [C.POINTER(abi.CloudView), C.POINTER(abi.CloudView), C.POINTER(abi.IcpParams),
                                    C.POINTER(C.c_double), C.POINTER(abi.IcpResult), C.POINTER(abi.IcpTrace),
                                    C.c_int, C.POINTER(C.c_double)]
        lib.orc_nn.restype = C.c_int
        lib.orc_nn.argtypes = [abi.CloudView, abi.CloudView, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        lib.orc_pca_features.restype = C.c_int
        lib.orc_pca_features.argtypes = [abi.CloudView, C.c_float, C.c_int, C.c_int, C.POINTER(abi.PcaOut)]
        lib.orc_num_threads.restype = C.c_int
        fpp = C.POINTER(C.c_float) * abi.NUM_CLASSES
        lib.orc_icp_run_trees.restype = C.c_int
        lib.orc_icp_run_trees.argtypes = [C.POINTER(abi.CloudView), C.POINTER(abi.CloudView), C.POINTER(abi.IcpParams),
                                          C.POINTER(C.c_double), C.POINTER(abi.IcpResult), fpp,
                                          C.POINTER(C.c_size_t)]
        lib.orc_map_update.restype = C.c_int
        lib.orc_map_update.argtypes = [C.POINTER(abi.CloudView), C.POINTER(C.c_double), C.POINTER(abi.CloudView),
                                       C.POINTER(C.c_double), C.POINTER(abi.CloudView), C.POINTER(abi.MapParams), fpp,
                                       C.POINTER(abi.MapInfo)]
        lib.orc_classify_nground.restype = C.c_int
        lib.orc_classify_nground.argtypes = [abi.CloudView, C.POINTER(abi.ClassifyParams), C.POINTER(abi.ClassifyOut)]
        lib.orc_fast_ground_filter.restype = C.c_int
        lib.orc_fast_ground_filter.argtypes = [abi.CloudView, C.POINTER(abi.GroundParams), C.POINTER(abi.GroundOut)]
        lib.orc_voxel_downsample.restype = C.c_int
        lib.orc_voxel_downsample.argtypes = [abi.CloudView, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_size_t)]
        lib.orc_sac_plane.restype = C.c_int
        lib.orc_sac_plane.argtypes = [abi.CloudView, C.c_double, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_float)]
        _LIB = lib
    return _LIB


def views(clouds):
    arr = (abi.CloudView * abi.NUM_CLASSES)()
    for c in range(abi.NUM_CLASSES):
        arr[c] = abi.cloud_view(clouds[c])
    return arr


def icp_run(tgt, src, params: abi.IcpParams, init_guess: np.ndarray, threads: int = 0, want_trace: bool = True,
            timings: np.ndarray | None = None):
    """tgt/src: six (n,12) float32 arrays. threads: 0 = reference-shaped (3 OpenMP sections), n>0 = n threads.
    Returns (result dict, trace dict or None)."""
    lib = load()
    res = abi.IcpResult()
    trace = abi.IcpTrace() if want_trace else None
    init = np.ascontiguousarray(init_guess, dtype=np.float64).reshape(16)
    tptr = timings.ctypes.data_as(C.POINTER(C.c_double)) if timings is not None else None
    lib.orc_icp_run(views(tgt), views(src), C.byref(params), init.ctypes.data_as(C.POINTER(C.c_double)),
                    C.byref(res), C.byref(trace) if trace is not None else None, threads, tptr)
    return abi.result_to_dict(res), (abi.trace_to_dict(trace) if trace is not None else None)


def nn(tgt: np.ndarray, src: np.ndarray, max_dist: float):
    lib = load()
    idx = np.empty(src.shape[0], dtype=np.int32)
    d2 = np.empty(src.shape[0], dtype=np.float32)
    lib.orc_nn(abi.cloud_view(tgt), abi.cloud_view(src), float(max_dist),
               idx.ctypes.data_as(C.POINTER(C.c_int32)), d2.ctypes.data_as(C.POINTER(C.c_float)))
    return idx, d2


def pca_features(cloud: np.ndarray, radius: float, k: int, stride: int = 1):
    lib = load()
    n = cloud.shape[0]
    ev = np.zeros((n, 3), np.float32)
    pr = np.zeros((n, 3), np.float32)
    nr = np.zeros((n, 3), np.float32)
    cnt = np.zeros(n, np.int32)
    out = abi.PcaOut(ev.ctypes.data_as(C.POINTER(C.c_float)), pr.ctypes.data_as(C.POINTER(C.c_float)),
                     nr.ctypes.data_as(C.POINTER(C.c_float)), cnt.ctypes.data_as(C.POINTER(C.c_int32)))
    lib.orc_pca_features(abi.cloud_view(cloud), float(radius), int(k), int(stride), C.byref(out))
    return {"eigenvalues": ev, "principal": pr, "normal": nr, "pt_num": cnt}


def num_threads() -> int:
    return int(load().orc_num_threads())


def icp_run_trees(tgt, src, params: abi.IcpParams, init_guess: np.ndarray):
    """One registration (reference-shaped) that also returns the six clouds block1's kd-trees were built on
    (cregistration.hpp:1209-1232): (result dict, [six (n,12) arrays])."""
    lib = load()
    res = abi.IcpResult()
    init = np.ascontiguousarray(init_guess, dtype=np.float64).reshape(16)
    bufs = [np.zeros((max(t.shape[0], 1), 12), np.float32) for t in tgt]
    ptrs = (C.POINTER(C.c_float) * abi.NUM_CLASSES)(*[b.ctypes.data_as(C.POINTER(C.c_float)) for b in bufs])
    n = (C.c_size_t * abi.NUM_CLASSES)()
    lib.orc_icp_run_trees(views(tgt), views(src), C.byref(params), init.ctypes.data_as(C.POINTER(C.c_double)),
                          C.byref(res), ptrs, n)
    return abi.result_to_dict(res), [np.ascontiguousarray(bufs[c][: n[c]]) for c in range(abi.NUM_CLASSES)]


def map_update(map_clouds, map_pose, scan_down, scan_pose, params: abi.MapParams, trees=None):
    """MapManager::update_local_map on host clouds. Returns ([six (n,12) arrays], info dict)."""
    lib = load()
    mp = np.ascontiguousarray(map_pose, dtype=np.float64).reshape(16)
    sp = np.ascontiguousarray(scan_pose, dtype=np.float64).reshape(16)
    bufs = [np.zeros((max(map_clouds[c].shape[0] + scan_down[c].shape[0], 1), 12), np.float32)
            for c in range(abi.NUM_CLASSES)]
    ptrs = (C.POINTER(C.c_float) * abi.NUM_CLASSES)(*[b.ctypes.data_as(C.POINTER(C.c_float)) for b in bufs])
    info = abi.MapInfo()
    lib.orc_map_update(views(map_clouds), mp.ctypes.data_as(C.POINTER(C.c_double)), views(scan_down),
                       sp.ctypes.data_as(C.POINTER(C.c_double)), views(trees) if trees is not None else None,
                       C.byref(params), ptrs, C.byref(info))
    d = abi.map_info_to_dict(info)
    return [np.ascontiguousarray(bufs[c][: d["n"][c]]) for c in range(abi.NUM_CLASSES)], d


def classify_nground(cloud: np.ndarray, params: abi.ClassifyParams) -> dict:
    """CFilter::classify_nground_pts on host rows: {"pillar": (n,12), ..., "vertex": ..., "unground": ...}."""
    return abi.classify_call(load().orc_classify_nground, None, cloud, params)


def fast_ground_filter(cloud: np.ndarray, params: abi.GroundParams) -> dict:
    """CFilter::fast_ground_filter on host rows: {"ground", "ground_down", "unground"} as (n,12) rows."""
    return abi.ground_call(load().orc_fast_ground_filter, None, cloud, params)


def sac_plane(cloud: np.ndarray, threshold: float, max_iterations: int = 20):
    """pcl::SACSegmentation plane fit as plane_seg_ransac configures it: (found, refined inlier indices, coefficients)."""
    cloud = abi.as_aos48(cloud)
    inl = np.zeros(max(len(cloud), 1), np.int32)
    n = C.c_int32(0)
    coeff = np.zeros(4, np.float32)
    ok = load().orc_sac_plane(abi.cloud_view(cloud), float(threshold), int(max_iterations),
                              inl.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n), coeff.ctypes.data_as(C.POINTER(C.c_float)))
    return bool(ok), inl[: n.value].copy(), coeff


def voxel_downsample(cloud: np.ndarray, voxel_size: float) -> np.ndarray:
    """CFilter::voxel_downsample on host rows: one point per occupied voxel, in voxel-index order."""
    cloud = abi.as_aos48(cloud)
    out = np.zeros((max(len(cloud), 1), 12), np.float32)
    n = C.c_size_t(0)
    load().orc_voxel_downsample(abi.cloud_view(cloud), float(voxel_size), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n))
    return np.ascontiguousarray(out[: n.value])
