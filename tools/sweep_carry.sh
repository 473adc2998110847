# interior-point warm start carried across MPC steps (qp_warm_carry) on the bench workload: iterations and step times
for c in 0 1; do
  echo "== carry=$c"
  timeout 300 python tools/run_scp_once.py --batch 1024 --steps 10 --hp 10 --step-lo 4 --step-hi 7 --carry $c 2>&1 | grep -E "^step|total|TOT" | cut -c1-160
done
