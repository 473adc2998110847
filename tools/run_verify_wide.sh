mkdir -p gpurun_out/v3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v3/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v3/pytest_gpu.txt
timeout 400 python bench.py --hp 20 --batch 4096 --steps 10 --warmup 3 --skip-cpu > gpurun_out/v3/c3_hp20_b4096.json 2> gpurun_out/v3/c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --hp 50 --batch 148 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 4 --warmup 3 --skip-cpu --skip-assembly > gpurun_out/v3/c4_hp50_b148.json 2> gpurun_out/v3/c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --skip-cpu > gpurun_out/v3/bench.json 2> gpurun_out/v3/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ["c3_hp20_b4096","c4_hp50_b148","bench"]:
    d=json.loads([l for l in open(f"gpurun_out/v3/{f}.json") if l.startswith("{")][-1])
    print(f, round(d["value"]), d["ms_per_step"], d["roofline"]["frac"], d["stats"]["plan"], d["stats"]["status_counts_rank0"])
PY
