# usage: bash tools/run_scale.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out/scale
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale/bench_n$N.json 2> gpurun_out/scale/bench_n$N.err; echo "rc=$?"
cat gpurun_out/scale/bench_n$N.json
if [ "$N" = "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --batch 512 --skip-cpu > gpurun_out/scale/bench_n${N}_b512.json 2> gpurun_out/scale/bench_n${N}_b512.err; echo "rc=$?"
cat gpurun_out/scale/bench_n${N}_b512.json
fi
