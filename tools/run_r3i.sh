# Round 2 (second half), final check of the long-horizon path (left-looking factorisation + prefetching sweeps for the L2-resident factor)
D=gpurun_out/${1:-r3i}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"; tail -2 $D/pytest_parity.txt
timeout 2400 python -m pytest tests/test_gpu_workloads.py -m gpu -q -s > $D/pytest_workloads.txt 2>&1; echo "workloads rc=$?"; grep "^\[" $D/pytest_workloads.txt | cut -c1-250; tail -2 $D/pytest_workloads.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $D/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --hp 50 --batch 148 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 4 --warmup 3 --skip-cpu --skip-assembly > $D/c4_hp50_b148.json 2> $D/c4.err; echo "c4 rc=$?"
timeout 900 python bench.py --hp 50 --batch 1024 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 3 --warmup 3 --skip-cpu --skip-assembly > $D/c4_hp50_b1024.json 2> $D/c4b.err; echo "c4b rc=$?"
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
timeout 400 python bench.py --hp 20 --batch 4096 --steps 10 --warmup 3 --skip-cpu > $D/c3_hp20_b4096.json 2> $D/c3.err; echo "c3 rc=$?"
python - <<PY
import json
for f in ('c4_hp50_b148','c4_hp50_b1024','bench','c3_hp20_b4096'):
    try:
        d=json.load(open('$D/'+f+'.json')); print(f, 'value %.0f e2e %.0f ms/step %.1f frac %.4f rollout %.0f strong %.0f (frac %.4f) ipm/qp %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['rollout']['value'], d['north_star_strong']['value'], d['north_star_strong']['roofline_frac'], d['stats']['ipm_per_qp']), d['stats']['status_counts_rank0'])
    except Exception as e: print(f, 'FAILED', e)
PY
