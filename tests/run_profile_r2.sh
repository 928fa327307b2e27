#!/bin/bash
# Round-2 measurement session on one B200 (gpurun): GPU test-suite, one bench line per codec / config,
# the ncu launch list of the headline step and `ncu --set full` captures of the dominant kernels.
# usage: bash tests/run_profile_r2.sh <tag> [quick]
TAG=${1:-r2a}
QUICK=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
B() {  # name, args...
  n=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/bench_${n}_$TAG.json 2> gpurun_out/bench_${n}_$TAG.err; echo "bench $n rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${n}_$TAG.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("  %-10s %.1f us/step %.1f Gelem/s | enc %.1f us (frac %.3f) dec %.1f us (frac %.3f) search %s | e2e %.2f ms | cpu %s" % (
        "$n", d["ms_per_step"]*1e3, d["value"]/1e9, r["encode_ms"]*1e3, r["frac"], r["decode_ms"]*1e3, r["decode_frac"],
        r.get("search_ms"), d["e2e"]["ms_per_step"], (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print("  parse failed", e)
PY
}
B hsq --steps 50 --warmup 5
if [ -z "$QUICK" ]; then
B qsgd --codec qsgd --steps 50 --warmup 5
B terngrad --codec terngrad --steps 50 --warmup 5
B sign --codec sign --steps 50 --warmup 5
B topk --codec topk --steps 20 --warmup 5
B hsq_k12 --codec hsq --k-bit 12 --workload flat --steps 10 --warmup 3
B hsq_d8 --codec hsq --c-dim 8 --workload flat --steps 20 --warmup 3
B hsq_d32 --codec hsq --c-dim 32 --workload flat --steps 20 --warmup 3
fi
# launch list of the headline step (serialised, cold cache: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
# full captures: the encode kernel, the decode kernel
ncu --set full --clock-control none --import-source on -k regex:hsq_encode_tc2 -s 4 -c 2 -f -o gpurun_out/prof_enc_$TAG \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_enc_$TAG.log 2>&1; echo "ncu enc rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hsq_decode_reduce_staged -s 4 -c 2 -f -o gpurun_out/prof_dec_$TAG \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dec_$TAG.log 2>&1; echo "ncu dec rc=$?"
ls -la gpurun_out/*$TAG* | head -40
