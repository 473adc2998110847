D=gpurun_out/${1:-r2r}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"; tail -3 $D/pytest_parity.txt
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
python -c "
import json,sys
d=json.load(open('$D/bench.json')); print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp']); print('ROLLOUT', d['rollout']['value']); print('STRONG', d['north_star_strong']['value'])"
timeout 400 python bench.py --hp 20 --batch 4096 --steps 10 --warmup 3 --skip-cpu --skip-assembly > $D/c3_hp20_b4096.json 2> $D/c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --hp 50 --batch 148 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 4 --warmup 3 --skip-cpu --skip-assembly > $D/c4_hp50_b148.json 2> $D/c4.err; echo "c4 rc=$?"
python -c "
import json
for f in ('c3_hp20_b4096','c4_hp50_b148'):
    d=json.load(open('$D/'+f+'.json')); print(f, 'value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp'], 'rollout', d['rollout']['value'])"
bash tools/run_timers.sh $1 2>&1 | grep -v "^+" | grep -v rollout | head -24
