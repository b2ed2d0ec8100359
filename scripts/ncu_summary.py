"""Summarise ncu outputs into small text files for profiles/ (run here, no GPU needed)."""
import collections
import csv
import subprocess
import sys


def launch_summary(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    for i, r in enumerate(rows):
        if r[0] == "ID":
            hdr, data = r, rows[i + 1:]
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        name = r[ki].split("(")[0][:70]
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel device time from: ncu --metrics gpu__time_duration.sum --clock-control none ({path})\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'share_%':>8s} {'avg_us':>10s}\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:70s} {a[0]:8d} {a[1]:12.1f} {a[1] / tot * 100:8.1f} {a[1] / a[0]:10.1f}\n")
    print(open(out).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def raw_summary(rep, out):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# selected metrics from: ncu --set full --clock-control none --import-source on ({rep})\n")
        for r in data:
            f.write(f"\n== launch id {r[0]}: {r[kn][:60]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k:85s} {r[i]:>18s} {units[i]}\n")
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launch_summary(sys.argv[2], sys.argv[3])
    else:
        raw_summary(sys.argv[2], sys.argv[3])
