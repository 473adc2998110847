D=gpurun_out/${1:-r2p}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"; tail -3 $D/pytest_parity.txt
timeout 900 python -m pytest tests/test_gpu_workloads.py -m gpu -x -q -s -k "rollout" > $D/pytest_rollout.txt 2>&1; echo "rollout rc=$?"; tail -3 $D/pytest_rollout.txt
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
python -c "
import json,sys
d=json.load(open('$D/bench.json')); print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp']); print('ROLLOUT', d['rollout']); print('STRONG', d['north_star_strong']['value'])"
timeout 300 python bench.py --skip-cpu --skip-assembly --steps 50 > $D/bench50.json 2> $D/bench50.err; python -c "
import json,sys
d=json.load(open('$D/bench50.json')); print('BENCH50 value', d['value'], 'ROLLOUT', d['rollout']['value'])"
