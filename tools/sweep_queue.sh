# work-queue policy sweep on the 1024-instance benchmark workload: pinned (never parked) instances x QPs per pop
for pin in 0 111 222 333 444; do for q in 1 2 4; do
echo "== pinned $pin quantum $q"; SCPB200_PINNED=$pin SCPB200_QUANTUM=$q python tools/run_scp_once.py --batch 1024 --steps 12 --step-lo 4 --step-hi 7 | grep "^step" | awk '{ms+=$7; qp+=$10; ipm+=$13} END{printf "  solve ms total %.1f  QPs %d  ipm %d\n", ms, qp, ipm}'
done; done
