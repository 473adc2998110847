"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`)
per kernel: launches, total / average duration, share of all kernel time.
    python tools/ncu_launch_summary.py gpurun_out/launches_r01_final.csv "command that was profiled" > profiles/..._summary.txt"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"^void ", "", r[4]).split("(")[0]
    name = re.sub(r"\((?:bool|int)\)", "", name)
    a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
    a[0] += 1; a[1] += float(r[14]) * (1e-3 if r[13] == "ns" else 1.0)
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print("ncu --metrics gpu__time_duration.sum --clock-control none -c 400:", sys.argv[2])
print("(per-launch times under ncu are serialised and cold-cache; the SHARE per kernel is what is comparable with the bench line)\n")
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>10s}  block / grid")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {a[0]:8d} {a[1]:12.1f} {100 * a[1] / tot:6.2f}% {a[1] / a[0]:10.1f}  {a[2]} / {a[3]}")
