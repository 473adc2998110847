# Round 2 (second half): block-inverse triangular sweeps (SCP_BLK = 4) against the tile-wise sweeps (SCP_BLK = 1) and SCP_BLK = 2.
D=gpurun_out/${1:-r3a}
mkdir -p $D
P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
for v in blk4 blk1 blk2; do [ -x tools/microbench_chol_$v ] && timeout 120 tools/microbench_chol_$v > $D/microbench_chol_$v.txt 2>&1; done
grep "grid 444" $D/microbench_chol_*.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"; tail -3 $D/pytest_parity.txt
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
for v in blk1 blk2; do
  SCPB200_LIB=$P/libvariant_$v.so timeout 300 python bench.py --skip-cpu --skip-assembly > $D/bench_$v.json 2> $D/bench_$v.err; echo "bench $v rc=$?"
done
timeout 400 python bench.py --hp 20 --batch 4096 --steps 5 --warmup 3 --skip-cpu --skip-assembly > $D/c3.json 2> $D/c3.err; echo "c3 rc=$?"
SCPB200_LIB=$P/libvariant_blk1.so timeout 400 python bench.py --hp 20 --batch 4096 --steps 5 --warmup 3 --skip-cpu --skip-assembly > $D/c3_blk1.json 2> $D/c3_blk1.err; echo "c3 blk1 rc=$?"
python - <<PY
import json
for f in ('bench','bench_blk1','bench_blk2','c3','c3_blk1'):
    try:
        d=json.load(open('$D/'+f+'.json')); print(f, 'value %.0f e2e %.0f ms/step %.3f frac %.4f rollout %.0f strong %.0f ipm/qp %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['rollout']['value'], d['north_star_strong']['value'], d['stats']['ipm_per_qp']), d['stats']['status_counts_rank0'])
    except Exception as e: print(f, 'FAILED', e)
PY
timeout 1500 python -m pytest tests/test_gpu_workloads.py -m gpu -q -s -x > $D/pytest_workloads.txt 2>&1; echo "workloads rc=$?"; grep "^\[" $D/pytest_workloads.txt | cut -c1-260; tail -2 $D/pytest_workloads.txt
