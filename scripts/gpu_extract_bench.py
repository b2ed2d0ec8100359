"""CFilter::extract_semantic_pts on the GPU (voxel filter -> ground filter -> classification, chained in HBM) vs the CPU
restatements, stage by stage: wall clock around the C-ABI calls (host rows in, host clouds out)
and device time from the library's CUDA events."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from mulls_b200 import abi  # noqa: E402
from mulls_b200.registration import Context  # noqa: E402
from oracle import oracle  # noqa: E402
from test_ground import params, raw_scan  # noqa: E402


def med(fn, reps):
    ts, out = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3, out


ctx = Context(0, 1, 16, 200000)
raw, _ = raw_scan()
gp = params()
cp = abi.default_classify_params()
cp.neighbor_searching_radius, cp.neighbor_k, cp.neigh_k_min, cp.pca_down_rate = 0.7, 25, 7, 2
cp.fixed_num_downsampling, cp.unground_down_fixed_num, cp.random_seed = 1, 20000, 3
if len(sys.argv) > 1 and sys.argv[1] == "--profile":  # three calls for an ncu launch list
    for _ in range(3):
        ctx.extract_semantic_pts(raw, 0.05, gp, cp)
    sys.exit(0)
ctx.extract_semantic_pts(raw, 0.05, gp, cp)
dev = []
t_vox, down = med(lambda: (ctx.voxel_downsample(raw, 0.05), dev.append(ctx.stats()["ms_total"]))[0], 10)
d_vox = np.median(dev); dev.clear()
t_gf, g = med(lambda: (ctx.fast_ground_filter(down, gp), dev.append(ctx.stats()["ms_total"]))[0], 10)
d_gf = np.median(dev); dev.clear()
t_all, e = med(lambda: (ctx.extract_semantic_pts(raw, 0.05, gp, cp), dev.append(ctx.stats()["ms_total"]))[0], 10)
d_all = np.median(dev)
o_vox, od = med(lambda: oracle.voxel_downsample(raw, 0.05), 3)
o_gf, og = med(lambda: oracle.fast_ground_filter(od, gp), 3)
o_cls, oc = med(lambda: oracle.classify_nground(og["unground"], cp), 3)
same = np.array_equal(e["down"].view(np.uint32), od.view(np.uint32)) and all(
    np.array_equal(e[k].view(np.uint32), og[k].view(np.uint32)) for k in ("ground", "ground_down")) and all(
    np.array_equal(e[k].view(np.uint32), oc[k].view(np.uint32)) for k in abi.OUT_NAMES)
print(f"raw scan {raw.shape[0]} pts -> voxel 0.05 m: {down.shape[0]} pts; ground {g['ground'].shape[0]} / "
      f"{g['ground_down'].shape[0]}, unground {g['unground'].shape[0]}")
print(f"voxel_downsample: GPU {t_vox:.2f} ms wall ({d_vox:.2f} ms device) vs CPU restatement {o_vox:.1f} ms")
print(f"fast_ground_filter (RANSAC plane per cell): GPU {t_gf:.2f} ms wall ({d_gf:.2f} ms device) vs CPU restatement {o_gf:.1f} ms")
print(f"extract_semantic_pts (voxel + ground + classification, chained in HBM): GPU {t_all:.2f} ms wall ({d_all:.2f} ms device) "
      f"vs CPU restatements {o_vox + o_gf + o_cls:.1f} ms ({oracle.num_threads()} threads for the PCA); identical={same}")
print({k: int(v.shape[0]) for k, v in e.items()})
