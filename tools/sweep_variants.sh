#!/bin/bash
# Tuning sweep: run the bench workload (3 closed-loop steps) with each built variant of the library.
P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
run() { # lib threads ctas_cap batch
  echo "== lib=$1 threads=$2 cap=$3 batch=$4"
  SCPB200_LIB=$P/$1 SCPB200_THREADS=$2 SCPB200_CTAS_PER_SM=$3 python tools/run_scp_once.py --batch $4 --steps 3 --step-lo 4 --step-hi 7 2>&1 | grep -E "plan|step 2"
}
run libscpb200.so 256 1 148
run libscpb200.so 256 2 296
run libscpb200.so 128 2 296
run libscpb200.so 256 2 1024
run libscpb200.so 128 2 1024
run libvariant_512_1.so 512 1 148
run libvariant_512_1.so 512 1 1024
run libvariant_256_3.so 256 3 444
run libvariant_256_3.so 256 3 1024
run libvariant_128_3.so 128 3 444
run libvariant_128_3.so 128 3 1024
run libvariant_128_4.so 128 4 592
run libvariant_128_4.so 128 4 1024
