"""PCA feature kernel timed on B200 next to the CPU oracle: the reference's operating point is
<= 20-40k non-ground points, K = 25-50 neighbours within R = 0.6-1.0 m (script/config/*.txt)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulls_b200 import synth, abi
from mulls_b200.registration import Context
from oracle import oracle

pair = synth.make_pair(1000, "c2")
cloud = np.ascontiguousarray(np.concatenate([pair["tgt"][c] for c in (abi.FACADE, abi.PILLAR, abi.BEAM, abi.ROOF)], axis=0)[:40000])
ctx = Context(0, 1, 16, len(cloud) + 16)
for radius, k, stride in ((1.0, 50, 1), (0.6, 25, 1), (1.0, 50, 2)):
    for _ in range(2): g = ctx.pca_features(cloud, radius, k, stride)
    t0 = time.perf_counter(); n = 10
    for _ in range(n): g = ctx.pca_features(cloud, radius, k, stride)
    gpu_ms = (time.perf_counter() - t0) / n * 1e3
    t0 = time.perf_counter(); o = oracle.pca_features(cloud, radius, k, stride); cpu_ms = (time.perf_counter() - t0) * 1e3
    same = np.array_equal(g["pt_num"], o["pt_num"])
    nq = int((o["pt_num"] > 0).sum())
    # algorithmic bytes: every query reads its <= k neighbours' positions (16 B each) + writes 40 B of features
    alg = nq * (min(k, 64) * 16 + 40)
    print(f"n={len(cloud)} R={radius} k={k} stride={stride}: GPU call {gpu_ms:.2f} ms ({nq/gpu_ms*1e-3:.2f} M queries/s, "
          f"{alg/gpu_ms/1e6:.1f} GB/s algorithmic incl. ingest+D2H) vs oracle ({oracle.num_threads()} threads) {cpu_ms:.0f} ms; "
          f"counts equal={same}", flush=True)
