set -x
D=gpurun_out/${1:-r2h}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"
timeout 900 python -m pytest tests/test_gpu_workloads.py -m gpu -x -q -s -k "rollout" > $D/pytest_rollout.txt 2>&1; echo "rollout rc=$?"
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
tail -15 $D/pytest_parity.txt; tail -30 $D/pytest_rollout.txt; tail -5 $D/bench.err; python -c "
import json,sys
d=json.load(open('$D/bench.json')); print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp'], 'solve share', d['roofline']['solve_share_of_step']); print('ROLLOUT', d['rollout'])"
