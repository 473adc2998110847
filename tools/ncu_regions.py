"""Attribute warp-stall samples of an ncu source page (cuda,sass csv) to code regions, separating barrier waits
(time other warps spend waiting for the region that runs before the barrier) from issue-side samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
cur, hdr, out = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and cur and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = {}
        for k, v in zip(hdr, r): d.setdefault(k, v)
        try: s = int(d["# Samples"]); ie = int(d["Instructions Executed"])
        except ValueError: continue
        bar = int(d.get("stall_barrier", "0") or 0)
        out.append((cur, int(r[0]), s, bar, ie))
REG = {
 "ipm_core.cuh": [(56,140,"tile_potrf_inv"),(141,205,"chol_tiles (trsm+gemm)"),(206,262,"chol_solve_tiles"),(263,330,"ipm init/start point"),(331,420,"residual phase"),(421,432,"D + form call"),(433,470,"pass: w1/rhs"),(471,520,"pass: dz/ds/step"),(521,560,"update/exit")],
 "ops_pair.cuh": [(24,45,"mul_P"),(46,90,"response/mul_A"),(91,120,"forces"),(121,140,"add_At/coef2"),(141,180,"form: M + omega"),(181,225,"form: diag blocks/direct"),(226,270,"form: pair blocks (warp)")],
}
tot = sum(o[2] for o in out); toti = sum(o[4] for o in out)
agg = {}
for f, ln, s, bar, ie in out:
    name = f
    for lo, hi, nm in REG.get(f, []):
        if lo <= ln <= hi: name = f + ":" + nm; break
    a = agg.setdefault(name, [0, 0, 0]); a[0] += s - bar; a[1] += bar; a[2] += ie
print(f"{'region':55s} {'non-barrier':>11s} {'barrier':>8s} {'instr':>7s}")
for k, (nb, b, ie) in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
    print(f"{k:55s} {100*nb/tot:10.1f}% {100*b/tot:7.1f}% {100*ie/toti:6.1f}%")
