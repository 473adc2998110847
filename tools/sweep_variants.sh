#!/bin/bash
# Tuning sweep: run the bench workload (closed-loop steps) with each built variant of the library / launch shape.
# A batch of 8192 is throughput-bound (no step ends on one instance's chain), so it ranks launch shapes by throughput.
P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
B=${1:-8192}
run() { # lib threads max_ctas batch
  echo "== lib=$1 threads=$2 max_ctas=$3 batch=$4"
  SCPB200_LIB=$P/$1 SCPB200_THREADS=$2 SCPB200_MAX_CTAS=$3 python tools/run_scp_once.py --batch $4 --steps 5 --step-lo 4 --step-hi 7 2>&1 | grep -E "plan|step [234]" | cut -c1-150
}
run libscpb200.so 256 2 $B
run libscpb200.so 128 2 $B
run libscpb200.so 128 3 $B
run libscpb200.so 192 3 $B
run libvariant_192_3.so 192 3 $B
run libvariant_256_3.so 256 3 $B
