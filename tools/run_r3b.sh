# Round 2 (second half): pipelined triangular sweeps (operands prefetched across the barrier, shuffles inside the 8 critical lanes)
# against the phase version (-DSCP_SOLVE_PIPELINED=0), same box.
D=gpurun_out/${1:-r3b}
mkdir -p $D
P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
for v in pipe phase; do [ -x tools/microbench_chol_$v ] && timeout 120 tools/microbench_chol_$v > $D/microbench_chol_$v.txt 2>&1; done
grep "smem 76384" $D/microbench_chol_*.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"; tail -3 $D/pytest_parity.txt
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
SCPB200_LIB=$P/libvariant_phase.so timeout 300 python bench.py --skip-cpu --skip-assembly > $D/bench_phase.json 2> $D/bench_phase.err; echo "bench phase rc=$?"
timeout 300 python bench.py --skip-cpu --skip-assembly > $D/bench2.json 2> $D/bench2.err; echo "bench2 rc=$?"
timeout 400 python bench.py --hp 20 --batch 4096 --steps 5 --warmup 3 --skip-cpu --skip-assembly > $D/c3.json 2> $D/c3.err; echo "c3 rc=$?"
SCPB200_LIB=$P/libvariant_phase.so timeout 400 python bench.py --hp 20 --batch 4096 --steps 5 --warmup 3 --skip-cpu --skip-assembly > $D/c3_phase.json 2> $D/c3_phase.err; echo "c3 phase rc=$?"
python - <<PY
import json
for f in ('bench','bench_phase','bench2','c3','c3_phase'):
    try:
        d=json.load(open('$D/'+f+'.json')); print(f, 'value %.0f e2e %.0f ms/step %.3f frac %.4f rollout %.0f strong %.0f ipm/qp %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['rollout']['value'], d['north_star_strong']['value'], d['stats']['ipm_per_qp']), d['stats']['status_counts_rank0'], d['e2e'])
    except Exception as e: print(f, 'FAILED', e)
PY
timeout 1500 python -m pytest tests/test_gpu_workloads.py -m gpu -q -s > $D/pytest_workloads.txt 2>&1; echo "workloads rc=$?"; grep "^\[" $D/pytest_workloads.txt | cut -c1-260; tail -2 $D/pytest_workloads.txt
