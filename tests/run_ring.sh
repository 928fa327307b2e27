#!/bin/bash
# Multi-GPU ring session (gpurun --gpus N): parity, then the ring bench pipelined over stages vs as one chain.
N=${1:-4}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29521 tests/dist_check.py > gpurun_out/dist_check_ring_n$N.log 2>&1; echo "dist_check rc=$?"; grep -E "ring.*(OK|MISMATCH)|Error|error" gpurun_out/dist_check_ring_n$N.log | head
GQ_DIST_SHAPES=resnet50 GQ_DIST_CASES=ring:hsq timeout 600 $RUN --master-port 29522 tests/dist_check.py > gpurun_out/dist_check_ring_full_n$N.log 2>&1; echo "dist_check resnet50 staged rc=$?"; grep -E "ring.*(OK|MISMATCH)|Error|error" gpurun_out/dist_check_ring_full_n$N.log | head
GQ_P2P=0 GQ_RING_PARTS=1 timeout 600 $RUN --master-port 29539 bench.py --gpus $N --mode ring --steps 30 --warmup 5 > gpurun_out/bench_ring_nccl_n$N.json 2> gpurun_out/bench_ring_nccl_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_ring_nccl_n$N.json").read().strip().splitlines()[-1])
    print("ring N=%d NCCL send/recv: %.1f us/step  %.1f Gelem/s" % (d["n_gpus"], d["ms_per_step"]*1e3, d["value"]/1e9))
except Exception as e:
    print("bench parse failed", e)
PY
for parts in 4 1 8; do
  GQ_RING_PARTS=$parts timeout 600 $RUN --master-port 2953$parts bench.py --gpus $N --mode ring --steps 30 --warmup 5 > gpurun_out/bench_ring_p${parts}_n$N.json 2> gpurun_out/bench_ring_p${parts}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_ring_p${parts}_n$N.json").read().strip().splitlines()[-1])
    print("ring N=%d stages=$parts: %.1f us/step  %.1f Gelem/s  e2e %.2f ms  [%s]" % (d["n_gpus"], d["ms_per_step"]*1e3, d["value"]/1e9, d["e2e"]["ms_per_step"], d["config"]["exchange"]))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_ring_p${parts}_n$N.err").read()[-600:])
PY
done
