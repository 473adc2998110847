D=gpurun_out/${1:-r2q}
mkdir -p $D
timeout 400 python bench.py --hp 20 --batch 4096 --steps 10 --warmup 3 --skip-cpu > $D/c3_hp20_b4096.json 2> $D/c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --hp 50 --batch 148 --trust-radius-frac 0.2 --max-scp-iter 100 --steps 4 --warmup 3 --skip-cpu --skip-assembly > $D/c4_hp50_b148.json 2> $D/c4.err; echo "c4 rc=$?"
python -c "
import json
for f in ('c3_hp20_b4096','c4_hp50_b148'):
    d=json.load(open('$D/'+f+'.json')); print(f, 'value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['stats']['status_counts_rank0'], 'ipm/qp', d['stats']['ipm_per_qp'], d['stats']['plan'], 'rollout', d['rollout']['value'], 'asm', (d.get('roofline_assembly') or {}).get('frac'))"
P=$PWD/senquential-convex-programming-for-trajectory-planning_b200
SCPB200_LIB=$P/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 148 --steps 2 --hp 50 --step-lo 4 --step-hi 7 > $D/timers_hp50_512.txt 2>&1
SCPB200_LIB=$P/libvariant_timers.so timeout 300 python tools/run_scp_once.py --batch 1184 --steps 3 --hp 20 --step-lo 4 --step-hi 7 > $D/timers_hp20_512.txt 2>&1
cat $D/timers_hp50_512.txt $D/timers_hp20_512.txt
