#!/bin/bash
# End-of-round multi-GPU sanity session (gpurun --gpus N): parity of ps and ring on the final tree, ps bench,
# ring bench, the reference arm under torchrun.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -E "OK|MISMATCH|Error|error" gpurun_out/dist_check_n$N.log | head -20
timeout 600 $RUN --master-port 29512 tests/dist_check_full.py --steps 4 > gpurun_out/dist_full_n$N.log 2>&1; echo "dist_check_full rc=$?"; grep -E "OK|MISMATCH|Error|error" gpurun_out/dist_full_n$N.log | head -8
GQ_DIST_SHAPES=resnet50 GQ_DIST_CASES=ring:hsq timeout 600 $RUN --master-port 29522 tests/dist_check.py > gpurun_out/dist_check_ring_full_n$N.log 2>&1; echo "ring resnet50 rc=$?"; grep -E "OK|MISMATCH|Error|error" gpurun_out/dist_check_ring_full_n$N.log | head
for m in ps ring; do
timeout 600 $RUN --master-port 29514 bench.py --gpus $N --mode $m --steps 50 --warmup 5 > gpurun_out/bench_${m}_n$N.json 2> gpurun_out/bench_${m}_n$N.err; echo "bench $m rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${m}_n$N.json").read().strip().splitlines()[-1])
    print("N=%d $m: %.1f us/step  %.1f Gelem/s  encode %.1f decode %.1f  e2e %.2f ms  [%s]" % (d["n_gpus"], d["ms_per_step"]*1e3, d["value"]/1e9, d["roofline"]["encode_ms"]*1e3, d["roofline"]["decode_ms"]*1e3, d["e2e"]["ms_per_step"], d["config"]["exchange"][:90]))
except Exception as e:
    print("bench parse failed", e)
PY
done
timeout 600 $RUN --master-port 29516 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "reference arm rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_ref_n$N.json").read().strip().splitlines()[-1])
    print("reference arm N=%d: %.4f Gelem/s  cores %s  kind %s" % (d["n_gpus"], d["value"]/1e9, d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"]))
except Exception as e:
    print("ref parse failed", e)
PY
