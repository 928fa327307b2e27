#!/bin/bash
# A/B of the tcgen05 encode variants on the GPU box (each variant in its own process, bounded).
mkdir -p gpurun_out
LOG=gpurun_out/tc2_ab.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
for v in "$@"; do
  echo "=== $v" >> $LOG
  timeout 300 python tests/tc2_variants.py "$v" >> $LOG 2>&1
  echo "exit $?" >> $LOG
done
grep -E "^VARIANT|MISMATCH|exit [1-9]|Error|error" $LOG
