#!/bin/bash
# Perf-only subset of the GEMM matrix (shapes of the adapter GEMMs at config 2).
mkdir -p gpurun_out
LOG=gpurun_out/gemm_perf.log
: > $LOG
run() { timeout 90 python tools/gemm_probe.py "$@" >> $LOG 2>&1; rc=$?; if [ $rc -ne 0 ]; then echo "rc=$rc args=$*" >> $LOG; fi; }
for cg in ${CGS:-1 2}; do
  run $cg 0 0 8192 8192 8192
  run $cg 0 0 17184 2048 2560
  run $cg 0 0 17184 4096 2048
  run $cg 0 1 17184 2048 4096
  run $cg 1 1 4096 2048 17184
  run $cg 1 1 2048 2560 17184
done
cat $LOG
