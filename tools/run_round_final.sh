# End-of-round capture: GPU tests, smoke, both bench arms, the other BASELINE configs, ncu launch list and full captures of
# the dominant kernel (the default workload and the rollout entry) and of the assembly kernel.
D=gpurun_out/${1:-r2end}
mkdir -p $D
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $D/gpu.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"
timeout 2400 python -m pytest tests/test_gpu_workloads.py -m gpu -q -s > $D/pytest_workloads.txt 2>&1; echo "workloads rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $D/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $D/bench_reference.json 2> $D/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
timeout 400 python bench.py --hp 20 --batch 4096 --steps 10 --warmup 3 --skip-cpu > $D/c3_hp20_b4096.json 2> $D/c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --hp 50 --batch 148 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 4 --warmup 3 --skip-cpu --skip-assembly > $D/c4_hp50_b148.json 2> $D/c4.err; echo "c4 rc=$?"
for B in 1024 4096 16384 65536; do
  timeout 400 python bench.py --batch $B --steps 10 --warmup 3 --skip-cpu --skip-assembly 2> $D/c5_$B.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'n_gpus':1,'batch':$B,'e2e':d['e2e']['value'],'frac':d['roofline']['frac'],'rollout':d['rollout']['value'],'rollout_frac':d['rollout']['roofline_frac'],'ipm_per_qp':d['stats']['ipm_per_qp'],'p50_ms':d['stats']['p50_ms_per_mpc_step']}))" >> $D/c5_sweep_batch_n1.jsonl
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu > $D/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scp_solve -s 3 -c 1 -o $D/prof_scp -f python bench.py --steps 2 --warmup 3 --skip-cpu --skip-assembly > $D/ncu_full.log 2>&1; echo "ncu2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 3 -c 1 -o $D/prof_asm -f python bench.py --steps 1 --warmup 3 --skip-cpu > $D/ncu_asm.log 2>&1; echo "ncu3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mpc_setup -s 3 -c 1 -o $D/prof_k1 -f python bench.py --steps 1 --warmup 3 --skip-cpu --skip-assembly > $D/ncu_k1.log 2>&1; echo "ncu4 rc=$?"
tail -3 $D/pytest_parity.txt; tail -3 $D/pytest_workloads.txt; tail -2 $D/smoke.txt; cat $D/c5_sweep_batch_n1.jsonl
python -c "
import json
for f in ('bench','c3_hp20_b4096','c4_hp50_b148'):
    d=json.load(open('$D/'+f+'.json')); print(f, 'value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'rollout', d['rollout']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
r=json.load(open('$D/bench_reference.json')); print('REF', r['value'], r['config']['batch_per_gpu'], r['cpu_baseline']['cores'])"
