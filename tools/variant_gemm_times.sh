#!/bin/bash
# Per-GEMM times of build variants of the library (P2T_LIB_PATH), one bench.py process each, on the same box:
#   tools/variant_gemm_times.sh base variants/libp2t_x.so ... ; "base" = the in-tree library
for v in "$@"; do
  if [ "$v" = base ]; then unset P2T_LIB_PATH; else export P2T_LIB_PATH=$PWD/prot2text-v2-esm3_b200/$v; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-optimizer --no-stages 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms/step', round(d['ms_per_step'],4), 'gemms', d['roofline']['per_gemm_us'], 'sum', round(sum(d['roofline']['per_gemm_us']),1), 'mhz', d['clocks']['sm_mhz'])"
done
