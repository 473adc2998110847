set -x
D=gpurun_out/${1:-ncu}
mkdir -p $D
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scp_solve -s 3 -c 1 -o $D/prof_scp -f python bench.py --steps 2 --warmup 3 --skip-cpu --skip-assembly > $D/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la $D
