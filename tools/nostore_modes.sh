#!/bin/bash
# forward-GEMM times with the epilogue's stores modified (timing experiments; results are wrong in modes 1-3):
#   0 normal, 1 no stores at all (direct path, math only), 2 staged TMA stores into a 256-row window (L2 only),
#   3 staging (STS + proxy fence) without the TMA store
for m in 0 1 2 3 0; do
  P2T_DEBUG_NOSTORE=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-optimizer --no-stages --no-graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nostore mode $m', 'gemms', d['roofline']['per_gemm_us'], 'mhz', d['clocks']['sm_mhz'])"
done
