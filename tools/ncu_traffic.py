"""Extract per-launch DRAM traffic and duration of each kernel from .ncu-rep captures into profiles/ncu_traffic.json.
    python tools/ncu_traffic.py gpurun_out/prof_scp_v5.ncu-rep gpurun_out/prof_asm_v3.ncu-rep"""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = re.sub(r"^void ", "", d["Kernel Name"]).split("(")[0].split("<")[0]
        val = lambda k: float(d[k]) * UNIT.get(u[k], 1)
        out[name] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                     "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                     "duration_us_under_ncu": val("gpu__time_duration.sum"), "capture": os.path.basename(rep),
                     "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size")}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
