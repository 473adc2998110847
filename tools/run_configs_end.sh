# BASELINE.json configs[2], [3], [4] on one GPU with the end-of-round kernels (parity configs; throughput for the record)
mkdir -p gpurun_out/cfg
timeout 400 python bench.py --hp 20 --batch 4096 --steps 10 --warmup 3 --skip-cpu > gpurun_out/cfg/c3_hp20_b4096.json 2> gpurun_out/cfg/c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --hp 50 --batch 148 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 4 --warmup 3 --skip-cpu --skip-assembly > gpurun_out/cfg/c4_hp50_b148.json 2> gpurun_out/cfg/c4.err; echo "c4 rc=$?"
for B in 1024 2048 4096 8192 16384 32768 65536; do
  timeout 400 python bench.py --batch $B --steps 10 --warmup 3 --skip-cpu --skip-assembly 2> gpurun_out/cfg/c5_$B.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'n_gpus':1,'batch':$B,'e2e':d['e2e']['value'],'frac':d['roofline']['frac'],'plan':d['stats']['plan'],'ipm_per_qp':d['stats']['ipm_per_qp'],'p50_ms':d['stats']['p50_ms_per_mpc_step']}))" >> gpurun_out/cfg/c5_sweep_batch_n1.jsonl
done
cat gpurun_out/cfg/c5_sweep_batch_n1.jsonl
