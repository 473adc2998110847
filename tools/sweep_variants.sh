#!/bin/bash
# Tuning sweep: run the bench workload (closed-loop steps) with each built variant of the library / launch shape.
P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
run() { # lib threads max_ctas batch
  echo "== lib=$1 threads=$2 max_ctas=$3 batch=$4"
  SCPB200_LIB=$P/$1 SCPB200_THREADS=$2 SCPB200_MAX_CTAS=$3 python tools/run_scp_once.py --batch $4 --steps 5 --step-lo 4 --step-hi 7 2>&1 | grep -E "plan|step [234]" | cut -c1-130
}
run libscpb200.so 256 2 1024
run libscpb200.so 128 2 1024
run libscpb200.so 128 3 1024
run libscpb200.so 128 4 1024
run libscpb200.so 96 4 1024
run libscpb200.so 64 4 1024
run libvariant_192_3.so 192 3 1024
