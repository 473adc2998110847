D=gpurun_out/${1:-r2n}
mkdir -p $D
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
python -c "
import json,sys
d=json.load(open('$D/bench.json')); print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp']); print('ROLLOUT', d['rollout']['value']); print('STRONG', d['north_star_strong']['value'])"
bash tools/run_timers.sh $1 2>&1 | grep -v "^+"
