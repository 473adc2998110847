set -x
D=gpurun_out/${1:-r2d}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
SCPB200_LIB=$PWD/senquential-convex-programming-for-trajectory-planning_b200/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 1024 --steps 6 --step-lo 4 --step-hi 7 > $D/timers_b1024.txt 2>&1; echo "timers rc=$?"
tail -5 $D/pytest_parity.txt; python -c "
import json,sys
d=json.load(open('$D/bench.json')); print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp'])"
tail -22 $D/timers_b1024.txt
if [ "$2" = "tol" ]; then timeout 1500 python tools/exp_tolerance.py 1024 > $D/exp_tolerance.txt 2>&1; cat $D/exp_tolerance.txt | grep "^=="; fi
