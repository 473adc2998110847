D=gpurun_out/${1:-r2t}
mkdir -p $D
timeout 300 python bench.py --skip-cpu --skip-assembly > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
python -c "
import json,sys
d=json.load(open('$D/bench.json')); print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'ROLLOUT', d['rollout']['value'], 'STRONG', d['north_star_strong']['value'])"
timeout 600 python bench.py --hp 50 --batch 148 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 4 --warmup 3 --skip-cpu --skip-assembly > $D/c4_hp50_b148.json 2> $D/c4.err; echo "c4 rc=$?"
python -c "
import json
for f in ('c4_hp50_b148',):
    d=json.load(open('$D/'+f+'.json')); print(f, 'value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp'], 'rollout', d['rollout']['value'])"
P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
SCPB200_LIB=$P/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 148 --steps 2 --hp 50 --step-lo 4 --step-hi 7 > $D/timers_hp50_512.txt 2>&1
tail -18 $D/timers_hp50_512.txt
timeout 600 python -m pytest tests/test_gpu_workloads.py -m gpu -x -q -s -k "hp50 or Hp50 or trust" > $D/pytest_hp50.txt 2>&1; echo "hp50 tests rc=$?"; tail -3 $D/pytest_hp50.txt
