# 256- vs 512-thread CTAs on the long-horizon shapes (one CTA per SM): closed-loop steps of the bench workload
mkdir -p gpurun_out/wide
run() { # threads hp batch steps
  echo "== threads=$1 Hp=$2 batch=$3"
  SCPB200_THREADS=$1 timeout 300 python tools/run_scp_once.py --batch $3 --steps $4 --hp $2 --step-lo 4 --step-hi 7 2>&1 | grep -E "plan|step|Error|error" | cut -c1-200
}
{
run 256 20 4096 3
run 512 20 4096 3
run 384 20 4096 3
run 256 50 148 2
run 512 50 148 2
run 256 20 148 3
run 512 20 148 3
} > gpurun_out/wide/sweep.txt 2>&1
cat gpurun_out/wide/sweep.txt
