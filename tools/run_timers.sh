D=gpurun_out/${1:-tim}
mkdir -p $D
SCPB200_LIB=$PWD/senquential-convex-programming-for-trajectory-planning_b200/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 1024 --steps 6 --step-lo 4 --step-hi 7 > $D/timers_b1024.txt 2>&1; echo "timers rc=$?"
tail -23 $D/timers_b1024.txt
SCPB200_LIB=$PWD/senquential-convex-programming-for-trajectory-planning_b200/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 1024 --steps 6 --step-lo 4 --step-hi 7 --rollout 1 > $D/timers_rollout_b1024.txt 2>&1; echo "timers rc=$?"
tail -18 $D/timers_rollout_b1024.txt
