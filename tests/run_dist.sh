#!/bin/bash
# Multi-GPU session (gpurun --gpus N): parity scripts, then the bench at N ranks (fused and unfused exchange).
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -E "OK|MISMATCH|Error|error|exchange" gpurun_out/dist_check_n$N.log | head -20
timeout 600 $RUN --master-port 29512 tests/dist_check_full.py --steps 4 > gpurun_out/dist_full_n$N.log 2>&1; echo "dist_check_full rc=$?"; grep -E "OK|MISMATCH|Error|error" gpurun_out/dist_full_n$N.log | head -20
GQ_P2P_FUSED=0 timeout 600 $RUN --master-port 29513 tests/dist_check.py > gpurun_out/dist_check_unfused_n$N.log 2>&1; echo "dist_check (unfused) rc=$?"; grep -E "MISMATCH|Error|error|exchange" gpurun_out/dist_check_unfused_n$N.log | head
timeout 600 $RUN --master-port 29514 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("N=%d fused: %.1f us/step  %.1f Gelem/s  encode %.1f decode %.1f  e2e %.2f ms  [%s]" % (d["n_gpus"], d["ms_per_step"]*1e3, d["value"]/1e9, d["roofline"]["encode_ms"]*1e3, d["roofline"]["decode_ms"]*1e3, d["e2e"]["ms_per_step"], d["config"]["exchange"]))
except Exception as e:
    print("bench parse failed", e)
PY
GQ_P2P_FUSED=0 timeout 600 $RUN --master-port 29515 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_unfused_n$N.json 2> gpurun_out/bench_unfused_n$N.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_unfused_n$N.json").read().strip().splitlines()[-1])
    print("N=%d unfused: %.1f us/step  %.1f Gelem/s  encode %.1f decode %.1f  [%s]" % (d["n_gpus"], d["ms_per_step"]*1e3, d["value"]/1e9, d["roofline"]["encode_ms"]*1e3, d["roofline"]["decode_ms"]*1e3, d["config"]["exchange"]))
except Exception as e:
    print("bench parse failed", e)
PY
