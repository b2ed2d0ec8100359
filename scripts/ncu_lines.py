"""Per-source-line share of executed warp instructions / stall samples of one captured launch of an .ncu-rep
(compile with -lineinfo, capture with --import-source on):  python scripts/ncu_lines.py <report> [launch index] [top N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = collections.defaultdict(lambda: [0, 0, 0, ""]); seen = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; seen[cur] += 1; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or cur is None or not r[0].isdigit() or seen[cur] != which + 1: continue
    try:
        a = int(r[hdr.index("Instructions Executed")] or 0); b = int(r[hdr.index("Thread Instructions Executed")] or 0); c = int(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    k = (cur.split("/")[-1], int(r[0])); agg[k][0] += a; agg[k][1] += b; agg[k][2] += c; agg[k][3] = r[1][:96]
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[2] for v in agg.values()) or 1
print(f"launch {which}: {tot} warp instructions, mean active lanes {sum(v[1] for v in agg.values()) / tot:.1f}, {tots} samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:16]:16s}:{k[1]:4d} inst {v[0] / tot * 100:5.1f}% lanes {v[1] / max(v[0], 1):5.1f} smp {v[2] / tots * 100:5.1f}%  {v[3]}")
