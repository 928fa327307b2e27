#!/bin/bash
# A/B of the mbarrier wait flavours of the tcgen05 encode (GQ_TC2_WAIT bit 0 epilogue, bit 1 MMA warp)
mkdir -p gpurun_out
for w in 0 1 2 3; do
  echo "=== GQ_TC2_WAIT=$w"
  GQ_TC2_WAIT=$w timeout 300 python tests/tc2_variants.py "g3,pair,fmask,f2,r1" 2>&1 | grep -E "^VARIANT|MISMATCH|rror"
  GQ_TC2_WAIT=$w timeout 120 python tests/tc2_trace.py > gpurun_out/trace_wait$w.log 2>&1; tail -19 gpurun_out/trace_wait$w.log
done
