#!/bin/bash
# Runs the GEMM bring-up matrix, one process per variant with a timeout. Output -> gpurun_out/gemm_probe.log
mkdir -p gpurun_out
LOG=gpurun_out/gemm_probe.log
: > $LOG
run() { timeout 90 python tools/gemm_probe.py "$@" >> $LOG 2>&1; rc=$?; if [ $rc -ne 0 ]; then echo "rc=$rc args=$*" >> $LOG; fi; }
for cg in 1 2; do
  run $cg 0 0 256 256 128
  run $cg 0 0 128 256 64 f32
  run $cg 0 0 1000 520 328
  run $cg 0 1 512 512 256
  run $cg 1 0 512 512 256
  run $cg 1 1 512 512 256
  run $cg 1 1 300 264 1000 f32
  run $cg 0 0 4096 4096 4096
  run $cg 0 0 8192 8192 8192
  run $cg 1 1 4096 2048 16384
  run $cg 0 1 16384 2048 4096
done
cat $LOG
