set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a/gpu.txt
timeout 1500 python -m pytest tests/test_gpu_workloads.py -m gpu -x -q -s > gpurun_out/r2a/pytest_workloads.txt 2>&1; echo "workloads rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2a/pytest_parity.txt 2>&1; echo "parity rc=$?"
timeout 300 python bench.py --skip-cpu > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
SCPB200_LIB=$PWD/senquential-convex-programming-for-trajectory-planning_b200/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 1024 --steps 6 --step-lo 4 --step-hi 7 > gpurun_out/r2a/timers_b1024.txt 2>&1; echo "timers rc=$?"
tail -30 gpurun_out/r2a/pytest_workloads.txt; tail -5 gpurun_out/r2a/pytest_parity.txt; cat gpurun_out/r2a/bench.json | head -c 1500; tail -25 gpurun_out/r2a/timers_b1024.txt
