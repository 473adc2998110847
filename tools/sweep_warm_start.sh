for wrg in 0.3 1 3; do for smi in 3 4 5 6; do
echo "== relgap $wrg min_iter $smi"; python tools/run_scp_once.py --batch 1024 --steps 12 --step-lo 4 --step-hi 7 --warm-relgap $wrg --snap-min-iter $smi | grep "^step" | awk '{ms+=$7; qp+=$10; ipm+=$13} END{printf "  solve ms total %.1f  QPs %d  ipm %d  ipm/QP %.2f\n", ms, qp, ipm, ipm/qp}'
done; done
