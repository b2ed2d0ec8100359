"""profiles/traffic.json: DRAM bytes per k_search launch (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum) next to
the ALGORITHMIC bytes of the same launches, for the benched configuration (64 resident C2 pairs, seeds 1000..1063).

    python scripts/make_traffic_json.py <report.ncu-rep> <total algorithmic MB of one run, as gpu_search_profile.py prints it> [pairs]

The capture holds the first iterations of one run (one k_search launch per iteration in the host launch loop). Their algorithmic
bytes are 28 * (active sources + targets of the running pairs): the active sources and the running pairs of every iteration come
from the CPU oracle on the same pairs (the GPU's counts equal the oracle's in every iteration — tests/test_gpu_parity.py), the
targets per pair (after the intersection filter) from the run's total."""
import csv
import json
import os
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def oracle_counts(seed):
    import numpy as np  # noqa: F401
    from mulls_b200 import synth
    from oracle import oracle

    pair = synth.make_pair(seed, "c2")
    res, tr = oracle.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"], threads=2)
    used = [pair["params"].used_feature_type[c:c + 1] == b"1" for c in range(6)]
    # sources that ENTER iteration i: the trace holds the sizes after iteration i's shrinking, the first are the clouds themselves
    sizes = [[len(pair["src"][c]) for c in range(6)]] + [list(tr["n_src"][i]) for i in range(res["iters"])]
    return res["iters"], [sum(n for n, u in zip(row, used) if u) for row in sizes[: res["iters"]]]


def main():
    rep, total_mb = sys.argv[1], float(sys.argv[2])
    n_pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[0], rows[2:]
    kn, rd, wr, tm = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    units = rows[1]
    launches = []
    for r in data:
        if "k_search" not in r[kn] and "k_keep" not in r[kn]:
            continue
        launches.append({"kernel": r[kn].split("(")[0].replace("void ", "").replace("mulls::", ""),
                         "dram_bytes": float(r[rd].replace(",", "")) * unit[units[rd]] + float(r[wr].replace(",", "")) * unit[units[wr]],
                         "ms": float(r[tm].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[tm]]})
    with ProcessPoolExecutor(min(32, n_pairs)) as ex:
        counts = list(ex.map(oracle_counts, [1000 + i for i in range(n_pairs)]))
    max_it = max(c[0] for c in counts)
    S = [sum(c[1][i] for c in counts if c[0] > i) for i in range(max_it)]
    R = [sum(1 for c in counts if c[0] > i) for i in range(max_it)]
    total = total_mb * 1e6
    t_bar = (total - 28.0 * sum(S)) / (28.0 * sum(R))  # targets of the enabled classes per pair, after the intersection filter
    for i, l in enumerate(launches):
        l["iteration"] = i
        l["algorithmic_bytes"] = 28.0 * (S[i] + t_bar * R[i]) if i < max_it else None
        l["ratio"] = l["dram_bytes"] / l["algorithmic_bytes"] if l["algorithmic_bytes"] else None
    cap_d = sum(l["dram_bytes"] for l in launches)
    cap_a = sum(l["algorithmic_bytes"] for l in launches if l["algorithmic_bytes"])
    out = {"source": os.path.basename(rep), "configuration": f"{n_pairs} resident C2 pairs in one context (the benched batch), host launch loop",
           "launches": launches, "targets_per_pair": t_bar,
           "k_search_dram_bytes_per_launch": cap_d / len(launches), "k_search_algorithmic_bytes_per_launch_same_launches": cap_a / len(launches),
           "ratio_same_launches": cap_d / cap_a,
           "note": "the bench's roofline.bytes_per_launch averages over ALL launches of a step (later ones hold few running pairs); "
                   "compare `traffic` with `traffic_algorithmic_same_launches`"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    for l in launches:
        print(l)
    print({k: out[k] for k in ("k_search_dram_bytes_per_launch", "k_search_algorithmic_bytes_per_launch_same_launches", "ratio_same_launches", "targets_per_pair")})


if __name__ == "__main__":
    main()
