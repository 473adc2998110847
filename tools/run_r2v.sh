D=gpurun_out/${1:-r2v}
mkdir -p $D
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q > $D/pytest_parity.txt 2>&1; echo "parity rc=$?"; tail -2 $D/pytest_parity.txt
timeout 900 python -m pytest tests/test_gpu_workloads.py -m gpu -q -k "rollout or known_answer or philox or config2" > $D/pytest_workloads_subset.txt 2>&1; echo "workloads subset rc=$?"; tail -2 $D/pytest_workloads_subset.txt
timeout 300 python bench.py --skip-cpu > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('$D/bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'rollout', d['rollout']['value'], 'strong', d['north_star_strong']['value'], d['sharding_bitwise_ok']['ok'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scp_solve -s 3 -c 1 -o $D/prof_scp -f python bench.py --steps 2 --warmup 3 --skip-cpu --skip-assembly > $D/ncu_full.log 2>&1; echo "ncu rc=$?"
