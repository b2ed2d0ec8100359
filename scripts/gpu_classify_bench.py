"""classify_nground_pts on the GPU vs the CPU restatement: wall clock around the C-ABI call
(host rows in, ten host clouds out) and device time from the library's CUDA events."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from mulls_b200 import abi  # noqa: E402
from mulls_b200.registration import Context  # noqa: E402
from oracle import oracle  # noqa: E402
from test_classify import kitti_params, unground_cloud  # noqa: E402

ctx = Context(0, 1, 16, 200000)
ung = unground_cloud()
if len(sys.argv) > 1 and sys.argv[1] == "--profile":  # two calls of the KITTI variant, for an ncu launch list
    for _ in range(2):
        ctx.classify_nground(ung, kitti_params())
    ctx.classify_nground(ung, kitti_params(pca_down_rate=1, unground_down_fixed_num=40000))
    sys.exit(0)
for name, p in (("kitti urban (20000 of %d pts, r=0.7 k=25 stride 2, NMS, fixed numbers)" % ung.shape[0], kitti_params()),
                ("dense (12000 pts, r=1.0 k=50 stride 1, NMS)", kitti_params(pca_down_rate=1, neighbor_searching_radius=1.0,
                                                                            neighbor_k=50, neigh_k_min=8,
                                                                            unground_down_fixed_num=12000)),
                ("40000 pts, r=0.7 k=25 stride 1, NMS", kitti_params(pca_down_rate=1, unground_down_fixed_num=40000))):
    g = ctx.classify_nground(ung, p)
    tg, td = [], []
    for _ in range(10):
        t0 = time.perf_counter()
        g = ctx.classify_nground(ung, p)
        tg.append(time.perf_counter() - t0)
        td.append(ctx.stats()["ms_total"])
    to = []
    for _ in range(3):
        t0 = time.perf_counter()
        o = oracle.classify_nground(ung, p)
        to.append(time.perf_counter() - t0)
    same = all(np.array_equal(g[k].view(np.uint32), o[k].view(np.uint32)) for k in abi.OUT_NAMES)
    sizes = {k: int(g[k].shape[0]) for k in abi.OUT_NAMES}
    print(f"{name}: GPU {np.median(tg)*1e3:.2f} ms wall ({np.median(td):.2f} ms device) vs CPU restatement "
          f"({oracle.num_threads()} threads for the PCA) {np.median(to)*1e3:.1f} ms; identical={same}; {sizes}", flush=True)
