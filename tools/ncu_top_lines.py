"""Summarise an ncu source page: top CUDA-C source lines by warp-stall samples.
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_top_lines.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and cur and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            s, ie = int(d["# Samples"]), int(d["Instructions Executed"])
        except ValueError:
            continue
        stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not" not in k and v.isdigit() and int(v)}
        conf = d.get("L1 Wavefronts Shared Excessive", "0")
        out.append((s, ie, cur, r[0], r[1].strip()[:100], stalls, conf))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print(f"total samples {tot}, total warp-instructions {toti}")
byfile = {}
for s, ie, f, *_ in out:
    a = byfile.setdefault(f, [0, 0]); a[0] += s; a[1] += ie
for f, (s, ie) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} samples {100*s/tot:5.1f}%  inst {100*ie/toti:5.1f}%")
out.sort(key=lambda o: -o[0])
for s, ie, f, ln, src, st, conf in out[:N]:
    top = ",".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/tot:5.1f}% s {100*ie/toti:5.1f}% i  {f}:{ln:>4s}  {src:100s} [{top}] xs={conf}")
