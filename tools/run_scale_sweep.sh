# usage: bash tools/run_scale_sweep.sh N B   (under gpurun --gpus N): BASELINE configs[4], scenario-sharded sweep: B instances per GPU on N GPUs
N=$1; B=$2
mkdir -p gpurun_out/scale
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --batch $B --skip-cpu --skip-assembly > gpurun_out/scale/sweep_n${N}_b$B.json 2> gpurun_out/scale/sweep_n${N}_b$B.err; echo "rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/scale/sweep_n${N}_b$B.json').read().strip().splitlines()[-1])
print(json.dumps({'n_gpus': d['n_gpus'], 'batch_per_gpu': $B, 'batch_total': $B*d['n_gpus'], 'value': d['value'], 'e2e': d['e2e']['value'], 'ms_per_step': d['ms_per_step'], 'frac': d['roofline']['frac'], 'rollout': d['rollout']['value'], 'sharding_bitwise_ok': d['sharding_bitwise_ok']['ok']}))"
