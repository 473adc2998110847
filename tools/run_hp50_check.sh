mkdir -p gpurun_out/hp50b
timeout 600 python -m pytest tests -m gpu -x -q -k "hp50 or hp20 or dense" > gpurun_out/hp50b/pytest.txt 2>&1; tail -3 gpurun_out/hp50b/pytest.txt
timeout 300 python tools/run_scp_once.py --batch 148 --steps 2 --hp 50 --step-lo 4 --step-hi 7 2>&1 | grep -E "plan|step" | cut -c1-200 | tee gpurun_out/hp50b/hp50.txt
timeout 300 python tools/run_scp_once.py --batch 1024 --steps 3 --hp 10 --step-lo 4 --step-hi 7 2>&1 | grep -E "plan|step" | cut -c1-200 | tee gpurun_out/hp50b/hp10.txt
