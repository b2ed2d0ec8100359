"""Source-sharded registration (BASELINE config 5): per-iteration exchange of the claim
table (min), the counts (sum) and the per-class normal-equation sums (sum).

* CPU: the contiguous sharding helper.
* GPU, one device: (a) world size 1 with an identity all-reduce must reproduce the unsharded run
  bit-for-bit; (b) TWO shards driven concurrently from two host threads on the same GPU, with an
  all-reduce that really combines the two ranks' device buffers — the full exchange logic without
  needing two GPUs. Both are checked against the oracle like every other parity test."""
import threading

import numpy as np
import pytest

from mulls_b200 import abi, synth
from mulls_b200.dist import shard_sources, tensor_from_ptr


def test_shard_sources_partition():
    rng = np.random.default_rng(0)
    clouds = [rng.normal(size=(n, 12)).astype(np.float32) for n in (1001, 0, 17, 5, 300, 2)]
    parts = [shard_sources(clouds, r, 4) for r in range(4)]
    for c in range(6):
        cat = np.concatenate([parts[r][0][c] for r in range(4)], axis=0)
        np.testing.assert_array_equal(cat, clouds[c])
        assert [parts[r][1][c] for r in range(4)] == sorted(parts[r][1][c] for r in range(4))
        assert all(parts[r][2][c] == len(clouds[c]) for r in range(4))


@pytest.mark.gpu
def test_world1_identity_allreduce_equals_unsharded(small_pair):
    from mulls_b200.registration import Context

    ctx = Context(0, 1, 100000, 100000)
    ref, rtr = ctx.run_batch([small_pair], want_trace=True)
    shards, base, glob = shard_sources(small_pair["src"], 0, 1)
    res, tr = ctx.run_sharded(dict(small_pair, src=shards), base, glob, lambda *a: 0, want_trace=True)
    assert res["code"] == ref[0]["code"] and res["iters"] == ref[0]["iters"]
    np.testing.assert_array_equal(res["T"], ref[0]["T"])
    np.testing.assert_array_equal(tr["n_corr"], rtr[0]["n_corr"])
    np.testing.assert_array_equal(tr["n_src"], rtr[0]["n_src"])
    ctx.close()


@pytest.mark.gpu
def test_two_shards_on_one_gpu_match_oracle(oracle_mod, small_pair):
    import torch

    from mulls_b200.registration import Context

    world = 2
    ctxs = [Context(0, 1, 100000, 100000) for _ in range(world)]
    barrier = threading.Barrier(world)
    slots = [None] * world
    dev = torch.device("cuda", 0)

    def make_hook(rank):
        def hook(ptr, count, dtype, op, stream):
            torch.cuda.ExternalStream(stream, device=dev).synchronize()
            slots[rank] = tensor_from_ptr(ptr, count, dtype, dev)
            barrier.wait()
            if rank == 0:
                stack = torch.stack([s.clone() for s in slots])
                red = stack.sum(0) if op == 0 else stack.min(0).values
                for s in slots:
                    s.copy_(red)
                torch.cuda.synchronize()
            barrier.wait()
            return 0

        return hook

    out = [None] * world

    def worker(rank):
        shards, base, glob = shard_sources(small_pair["src"], rank, world)
        out[rank] = ctxs[rank].run_sharded(dict(small_pair, src=shards), base, glob, make_hook(rank), want_trace=True)

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
        assert not t.is_alive()
    o, ot = oracle_mod.icp_run(small_pair["tgt"], small_pair["src"], small_pair["params"], small_pair["init_guess"])
    for r in range(world):
        g, gt = out[r]
        assert g["code"] == o["code"] and g["iters"] == o["iters"]
        np.testing.assert_array_equal(gt["n_corr"], ot["n_corr"])
        np.testing.assert_array_equal(gt["n_src"], ot["n_src"])
        dt, dr = synth.pose_error(g["T"], o["T"])
        assert dt <= 1e-4 and dr <= 1e-4
    np.testing.assert_array_equal(out[0][0]["T"], out[1][0]["T"])  # every rank holds the same result
    for c in ctxs:
        c.close()
