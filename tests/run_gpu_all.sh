#!/bin/bash
# One GPU-box session: kernel variants A/B + trace, then the GPU test-suite, then the bench lines.
mkdir -p gpurun_out
bash tests/run_tc2_ab.sh "$@"
python tests/tc2_trace.py "g3,pair,fmask,f2,r2" > gpurun_out/trace_r2.log 2>&1; tail -13 gpurun_out/trace_r2.log
python tests/tc2_trace.py "g3,pair,fmask,f2,r1" > gpurun_out/trace_r1.log 2>&1; tail -13 gpurun_out/trace_r1.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_hsq.json 2> gpurun_out/bench_hsq.err; tail -3 gpurun_out/bench_hsq.err; cat gpurun_out/bench_hsq.json
