P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
mkdir -p gpurun_out/hp50
SCPB200_LIB=$P/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 148 --steps 2 --hp 50 --step-lo 4 --step-hi 7 > gpurun_out/hp50/timers_hp50_512.txt 2>&1
SCPB200_LIB=$P/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 74 --steps 2 --hp 50 --step-lo 4 --step-hi 7 > gpurun_out/hp50/timers_hp50_512_b74.txt 2>&1
SCPB200_LIB=$P/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 1184 --steps 3 --hp 20 --step-lo 4 --step-hi 7 > gpurun_out/hp50/timers_hp20_512.txt 2>&1
cat gpurun_out/hp50/timers_hp50_512.txt gpurun_out/hp50/timers_hp50_512_b74.txt gpurun_out/hp50/timers_hp20_512.txt
