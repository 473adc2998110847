set -x
mkdir -p gpurun_out/r2b
timeout 2400 python -m pytest tests/test_gpu_workloads.py -m gpu -q -s > gpurun_out/r2b/pytest_workloads.txt 2>&1; echo "workloads rc=$?"
grep "^\[" gpurun_out/r2b/pytest_workloads.txt; tail -40 gpurun_out/r2b/pytest_workloads.txt
