"""On-disk formats at the boundary of the hot path: scans in, poses out.

Readers return C-contiguous (n, 12) float32 arrays in the pcl::PointXYZINormal row layout the C-ABI consumes
(x y z 1 | normal_x normal_y normal_z 0 | intensity curvature 0 0), optionally in pinned host memory so that the
H2D copy of `mulls_icp_run_batch` is a plain DMA. Semantics follow the reference's DataIo:

* read_pcd        <- DataIo::read_pcd_file            include/common/dataio.hpp:279-287 (pcl::io::loadPCDFile)
* read_kitti_bin  <- DataIo::read_bin_file            include/common/dataio.hpp:357-377
* read_cloud_block<- DataIo::read_pc_cloud_block      include/common/dataio.hpp:1732-1756
* write_lo_pose_* <- DataIo::write_lo_pose_overwrite / _append   include/common/dataio.hpp:1896-1926

The product's readers are the NATIVE ones of the C-ABI (mulls_scan_probe / mulls_scan_read / mulls_pose_write,
csrc/scan_io.h), wrapped here as read_cloud_block_native / write_lo_pose_native; the numpy readers above are the
independent second implementation the CPU tests compare them with.
"""
from __future__ import annotations

import numpy as np

_FIELD_COL = {"x": 0, "y": 1, "z": 2, "normal_x": 4, "normal_y": 5, "normal_z": 6, "intensity": 8, "curvature": 9}


def _alloc(n: int, pinned: bool) -> np.ndarray:
    if pinned:
        import torch

        out = torch.zeros((n, 12), dtype=torch.float32).pin_memory().numpy()  # the ndarray keeps the tensor alive
    else:
        out = np.zeros((n, 12), dtype=np.float32)
    out[:, 3] = 1.0
    return out



def read_pcd(path: str, pinned: bool = False) -> np.ndarray:
    """PCD v0.7, DATA binary or ascii, float32 fields among x y z intensity normal_x normal_y normal_z curvature."""
    with open(path, "rb") as f:
        fields, sizes, types, counts, npts, data = None, None, None, None, None, None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PCD header without DATA line")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok or tok[0].startswith("#"):
                continue
            key = tok[0].upper()
            if key == "FIELDS":
                fields = tok[1:]
            elif key == "SIZE":
                sizes = [int(v) for v in tok[1:]]
            elif key == "TYPE":
                types = tok[1:]
            elif key == "COUNT":
                counts = [int(v) for v in tok[1:]]
            elif key == "POINTS":
                npts = int(tok[1])
            elif key == "WIDTH" and npts is None:
                npts = int(tok[1])
            elif key == "DATA":
                data = tok[1].lower()
                break
        if fields is None or npts is None:
            raise ValueError("incomplete PCD header")
        counts = counts or [1] * len(fields)
        if any(s != 4 or t != "F" or c != 1 for s, t, c in zip(sizes, types, counts)):
            raise ValueError("only 4-byte float fields with COUNT 1 are supported")
        if data == "binary":
            raw = np.frombuffer(f.read(npts * 4 * len(fields)), dtype="<f4").reshape(npts, len(fields))
        elif data == "ascii":
            raw = np.loadtxt(f, dtype=np.float32, ndmin=2)[:npts]
        else:
            raise ValueError(f"unsupported PCD DATA {data!r} (binary_compressed is not handled)")
    out = _alloc(npts, pinned)
    for j, name in enumerate(fields):
        col = _FIELD_COL.get(name)
        if col is not None:
            out[:, col] = raw[:, j]
    return out


def read_kitti_bin(path: str, pinned: bool = False, reference_eof_point: bool = True) -> np.ndarray:
    """KITTI velodyne .bin (x y z reflectance float32). intensity = reflectance * 255 (dataio.hpp:372).
    The reference's read loop tests eof() only after the failed read, so it appends one default-constructed
    point (0,0,0, intensity 0) at the end; `reference_eof_point=True` reproduces that."""
    raw = np.fromfile(path, dtype="<f4")
    raw = raw[: (raw.size // 4) * 4].reshape(-1, 4)
    n = raw.shape[0] + (1 if reference_eof_point else 0)
    out = _alloc(n, pinned)
    out[: raw.shape[0], 0:3] = raw[:, 0:3]
    out[: raw.shape[0], 8] = raw[:, 3] * np.float32(255)
    return out


def cloud_bounds(cloud: np.ndarray) -> tuple:
    """CloudUtility::get_cloud_bbx (utility.hpp:817-848): min_x min_y min_z max_x max_y max_z as doubles."""
    if cloud.shape[0] == 0:
        big = 1.7976931348623157e308
        return (big, big, big, -big, -big, -big)
    xyz = cloud[:, 0:3].astype(np.float64)
    mn, mx = xyz.min(0), xyz.max(0)
    return (mn[0], mn[1], mn[2], mx[0], mx[1], mx[2])


def read_cloud_block(path: str, normalize_intensity: bool = False, pinned: bool = False) -> dict:
    """read_pc_cloud_block: raw cloud + local_bound (+ intensity rescaled to 0..255 in float, dataio.hpp:1738-1750)."""
    cloud = read_kitti_bin(path, pinned) if path.lower().endswith(".bin") else read_pcd(path, pinned)
    bound = cloud_bounds(cloud)
    if normalize_intensity and cloud.shape[0]:
        inten = cloud[:, 8]
        lo, hi = np.float32(inten.min()), np.float32(inten.max())
        scale = np.float32(255.0 / float(hi - lo))  # float intesnity_scale = 255.0 / (max - min)
        cloud[:, 8] = (inten - lo) * scale
    return {"pc_raw": cloud, "local_bound": bound}


def _pose_line(T: np.ndarray) -> str:
    T = np.asarray(T, dtype=np.float64)
    return " ".join("%.8g" % T[r, c] for r in range(3) for c in range(4)) + "\n"  # out << setprecision(8)


def write_lo_pose_overwrite(T: np.ndarray, path: str) -> bool:
    with open(path, "w") as f:
        f.write(_pose_line(T))
    return True


def write_lo_pose_append(T: np.ndarray, path: str) -> bool:
    with open(path, "a") as f:
        f.write(_pose_line(T))
    return True


class PinnedRows:
    """(n, 12) float32 rows in pinned host memory of the library (mulls_host_alloc); freed with the object."""

    def __init__(self, lib, n: int):
        import ctypes as C

        self._lib, self._ptr = lib, lib.mulls_host_alloc(max(n, 1) * 48)
        if not self._ptr:
            raise MemoryError("mulls_host_alloc failed (no CUDA device?)")
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(C.c_float)), shape=(max(n, 1), 12))[:n]

    def __del__(self):
        if getattr(self, "_ptr", None):
            self._lib.mulls_host_free(self._ptr)
            self._ptr = None


def read_cloud_block_native(path: str, normalize_intensity: bool = False, pinned: bool = False) -> dict:
    """DataIo::read_pc_cloud_block through the C-ABI: rows (optionally in pinned memory) + local_bound."""
    import ctypes as C

    from . import abi

    lib = abi.load_library()
    n = C.c_size_t(0)
    rc = lib.mulls_scan_probe(path.encode(), C.byref(n))
    if rc != 0:
        raise IOError(f"mulls_scan_probe({path!r}) failed: {rc}")
    keep = PinnedRows(lib, n.value) if pinned else None
    rows = keep.array if pinned else np.empty((n.value, 12), np.float32)
    bound = (C.c_double * 6)()
    got = C.c_size_t(0)
    rc = lib.mulls_scan_read(path.encode(), rows.ctypes.data_as(C.POINTER(C.c_float)), n.value, C.byref(got), bound,
                             1 if normalize_intensity else 0)
    if rc != 0:
        raise IOError(f"mulls_scan_read({path!r}) failed: {rc}")
    return {"pc_raw": rows[: got.value], "local_bound": tuple(bound), "_pinned": keep}


def write_lo_pose_native(T: np.ndarray, path: str, overwrite: bool = False) -> bool:
    import ctypes as C

    from . import abi

    pose = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    return abi.load_library().mulls_pose_write(path.encode(), pose.ctypes.data_as(C.POINTER(C.c_double)), 1 if overwrite else 0) == 0
