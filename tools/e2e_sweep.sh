#!/bin/bash
# e2e (host buffers -> HostStager -> step -> loss.item()) for several staging settings, same box:
#   tools/e2e_sweep.sh "P2T_STAGE_STREAMS=3" "P2T_STAGE_MODE=pull P2T_STAGE_PULL_CTAS=32" ...
for n in "$@"; do
  env $n python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-optimizer --no-stages 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$n', 'e2e pairs/s', round(e['value'],1), 'ms', round(e['ms_per_step'],4), 'GB/s', round(e['h2d_gb_per_s'],2), 'peak', round(e['h2d_peak_gb_per_s'],2), 'value', round(d['value'],1))"
done
