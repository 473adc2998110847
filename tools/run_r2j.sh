set -x
D=gpurun_out/${1:-r2j}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "rate_rows" > $D/pytest_rate.txt 2>&1; echo "rate rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"
timeout 600 python bench.py > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $D/bench_ref.json 2> $D/bench_ref.err; echo "ref rc=$?"
tail -25 $D/pytest_rate.txt; tail -8 $D/pytest_parity.txt; tail -5 $D/bench.err; python -c "
import json,sys
d=json.load(open('$D/bench.json')); print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp']); print('ROLLOUT', d['rollout']['value']); print('STRONG', d['north_star_strong']); print('SHARD', d['sharding_bitwise_ok']); print('SETUP', d['roofline_setup']); print('ASM', d['roofline_assembly']); print('CPU', d.get('cpu_baseline'))
r=json.load(open('$D/bench_ref.json')); print('REF', r['value'], r['config']['batch_per_gpu'], r['ms_per_step'])"
