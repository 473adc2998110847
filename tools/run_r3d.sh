# Round 2 (second half): left-looking factorisation for the L2-resident factor (Hp = 50).
D=gpurun_out/${1:-r3d}
mkdir -p $D
P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"; tail -2 $D/pytest_parity.txt
timeout 900 python -m pytest tests/test_gpu_workloads.py -m gpu -q -s -k "hp50 or Hp50 or trust or config4" > $D/pytest_hp50.txt 2>&1; echo "hp50 tests rc=$?"; grep "^\[" $D/pytest_hp50.txt | cut -c1-250; tail -2 $D/pytest_hp50.txt
timeout 600 python bench.py --hp 50 --batch 148 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 4 --warmup 3 --skip-cpu --skip-assembly > $D/c4_hp50_b148.json 2> $D/c4.err; echo "c4 rc=$?"
python - <<PY
import json
d=json.load(open('$D/c4_hp50_b148.json')); print('c4', 'value %.0f ms/step %.1f frac %.4f rollout %.0f strong %.0f (frac %.4f) ipm/qp %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['rollout']['value'], d['north_star_strong']['value'], d['north_star_strong']['roofline_frac'], d['stats']['ipm_per_qp']), d['stats']['status_counts_rank0'])
PY
SCPB200_LIB=$P/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 148 --steps 2 --hp 50 --step-lo 4 --step-hi 7 > $D/timers_hp50_512.txt 2>&1
tail -18 $D/timers_hp50_512.txt
