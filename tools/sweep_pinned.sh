for pn in 111 222 296 370 444; do
  for w in 5 3; do
    SCPB200_PINNED=$pn timeout 300 python bench.py --steps 20 --warmup $w --skip-cpu --skip-assembly 2>/dev/null | tail -1 > gpurun_out/b.json
    python - <<PY
import json
d=json.load(open('gpurun_out/b.json')); print('pinned $pn warmup $w: value %.0f e2e %.0f ms/step %.3f p50 %.3f strong %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['stats']['p50_ms_per_mpc_step'], d['north_star_strong']['value']))
PY
  done
done
