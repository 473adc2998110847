"""Instruction counts per kernel of the shipped library from `cuobjdump -sass` (evidence for the tensor / TMA paths):
    python tools/sass_counts.py [path/to/libscpb200.so] > profiles/rNN_sass_counts.txt
DMMA = mma.sync.m8n8k4.f64 (the Blackwell FP64 tensor path; FP64 has no tcgen05 form), UBLKCP = cp.async.bulk (TMA bulk
store of the assembly kernel), DFMA / DADD / DMUL = the FP64 pipe, LDS / STS = shared memory, BAR = CTA barriers."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "senquential-convex-programming-for-trajectory-planning_b200", "libscpb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
KEYS = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR", "ATOM", "total"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
        if op in ("RED", "ATOMG", "ATOMS"):
            counts[cur]["ATOM"] += 1
print(f"{os.path.basename(lib)}: architectures {arch}, {len(counts)} kernels")
print(f"{'kernel':78s} " + " ".join(f"{k:>7s}" for k in KEYS))
tot = collections.Counter()
for name, c in counts.items():
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\([^()]*\)$", "", dem).replace("(bool)", "").replace("(int)", "").replace("void ", "")
    print(f"{dem[:78]:78s} " + " ".join(f"{c[k]:7d}" for k in KEYS))
    tot.update(c)
print(f"{'all kernels':78s} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
