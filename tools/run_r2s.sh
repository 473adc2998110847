D=gpurun_out/${1:-r2s}
mkdir -p $D
for w in 8 10 12 14 16 18 20 24 30; do echo "== qp_warm_max_iter $w"; python tools/run_scp_once.py --batch 1024 --steps 12 --step-lo 4 --step-hi 7 --warm-max-iter $w 2>&1 | grep "^step" | awk '{ms+=$7; qp+=$10; it+=$13} END {printf "total solve ms %.2f  QPs %d  ipm %d  -> %.0f QP/s\n", ms, qp, it, qp/ms*1000}'; done 2>&1 | tee $D/sweep_wmi.txt
for B in 512 1024; do for parts in 1 2 3 4 6 8; do echo "== B $B parts $parts"; SCPB200_ASM_PARTS=$parts python tools/time_assemble.py --batch $B --reps 20 2>&1 | tail -1; done; done 2>&1 | tee $D/sweep_asm_parts.txt
