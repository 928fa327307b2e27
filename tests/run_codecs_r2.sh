#!/bin/bash
# GPU session: parity suite, then one bench line per elementwise codec (QSGD, TernGrad, sign, top-k)
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
B() {
  n=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/bench_${n}_$TAG.json 2> gpurun_out/bench_${n}_$TAG.err; echo "bench $n rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${n}_$TAG.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("  %-10s %.1f us/step %.1f Gelem/s | enc %.1f us (frac %.3f) dec %.1f us (frac %.3f) | e2e %.2f ms" % (
        "$n", d["ms_per_step"]*1e3, d["value"]/1e9, r["encode_ms"]*1e3, r["frac"], r["decode_ms"]*1e3, r["decode_frac"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("  parse failed", e)
PY
}
for c in ${CODECS:-qsgd terngrad sign topk}; do
  B $c --codec $c --steps 50 --warmup 5 --no-cpu-baseline
done
