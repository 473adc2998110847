# End-of-round capture: GPU tests, both bench arms, ncu launch list and one full capture of the dominant kernel.
set -x
mkdir -p gpurun_out/f3
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/f3/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f3/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f3/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f3/bench_reference.json 2> gpurun_out/f3/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/f3/bench.json 2> gpurun_out/f3/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f3/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/f3/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scp_solve -s 3 -c 1 -o gpurun_out/f3/prof_scp -f python bench.py --steps 2 --warmup 3 --skip-cpu --skip-assembly > gpurun_out/f3/ncu_full.log 2>&1; echo "ncu2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scp_solve -s 2 -c 1 -o gpurun_out/f3/prof_scp_hp20 -f python bench.py --hp 20 --batch 4096 --steps 1 --warmup 3 --skip-cpu --skip-assembly > gpurun_out/f3/ncu_full_hp20.log 2>&1; echo "ncu3 rc=$?"
tail -3 gpurun_out/f3/pytest_gpu.txt; cat gpurun_out/f3/smoke.txt | tail -2; cat gpurun_out/f3/bench.json; cat gpurun_out/f3/bench_reference.json
