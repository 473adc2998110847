# ncu --set full of the long-horizon kernel (Hp = 50, factor in the L2-resident workspace)
D=gpurun_out/${1:-r3e}
mkdir -p $D
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scp_solve -s 1 -c 1 -o $D/prof_hp50 -f python tools/run_scp_once.py --batch 148 --steps 2 --hp 50 --step-lo 4 --step-hi 7 --max-scp-iter 3 > $D/ncu_hp50.log 2>&1; echo "ncu rc=$?"; tail -3 $D/ncu_hp50.log
