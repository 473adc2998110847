D=gpurun_out/${1:-r3c}
mkdir -p $D
for v in pipe32 nowork nocrit empty; do timeout 120 tools/microbench_chol_p_$v > $D/mb_$v.txt 2>&1; echo "== $v"; grep "smem 76384" $D/mb_$v.txt; done
