P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
run() { echo "== threads=$1 max_ctas=$2 want_H=$3"; SCPB200_THREADS=$1 SCPB200_MAX_CTAS=$2 SCPB200_WANT_H=$3 python tools/run_scp_once.py --batch 1024 --steps 4 --step-lo 4 --step-hi 7 2>&1 | grep -E "plan|step [23]" | cut -c1-110; }
run 128 2 1
run 128 2 0
run 256 2 0
run 128 3 0
