for cfg in "256 2" "128 3" "96 3" "160 3" "192 3"; do set -- $cfg
echo "== threads=$1 ctas=$2"; SCPB200_THREADS=$1 SCPB200_MAX_CTAS=$2 python bench.py --skip-cpu --skip-assembly 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stats']['p50_ms_per_mpc_step'], d['stats']['plan'])"
done
