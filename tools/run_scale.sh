# usage: bash tools/run_scale.sh N   (under gpurun --gpus N): the driver's launch line; the JSON line carries the weak-scaling
# value, the north-star strong leg (4096 instances in total) and the on-hardware sharding bit-identity check
N=$1
mkdir -p gpurun_out/scale
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale/bench_n$N.json 2> gpurun_out/scale/bench_n$N.err; echo "rc=$?"
tail -3 gpurun_out/scale/bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/scale/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'rollout', d['rollout']['value'])
print('strong', d['north_star_strong'])
print('shard', d['sharding_bitwise_ok'])"
