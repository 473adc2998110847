set -x
D=gpurun_out/${1:-tests}
mkdir -p $D
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"
timeout 2400 python -m pytest tests/test_gpu_workloads.py -m gpu -q -s > $D/pytest_workloads.txt 2>&1; echo "workloads rc=$?"
tail -15 $D/pytest_parity.txt; grep -E "^\[|worst QP|passed|failed|Error|assert" $D/pytest_workloads.txt | head -60
