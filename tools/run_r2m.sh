D=gpurun_out/${1:-r2m}
mkdir -p $D
timeout 300 python tools/diag_rollout_generic.py > $D/diag_rollout.txt 2>&1; echo "diag rc=$?"; cat $D/diag_rollout.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"; tail -5 $D/pytest_parity.txt
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
python -c "
import json,sys
d=json.load(open('$D/bench.json')); print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp']); print('ROLLOUT', d['rollout']['value']); print('STRONG', d['north_star_strong']['value'])"
bash tools/run_timers.sh $1 2>&1 | grep -v "^+"
