#!/bin/bash
# 8-GPU ring session: parity at full size, then the ring bench over NCCL and peer-to-peer (default stages).
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
GQ_DIST_SHAPES=resnet50 GQ_DIST_CASES=ring:hsq,ps:hsq timeout 300 $RUN --master-port 29522 tests/dist_check.py > gpurun_out/dist_check_ring_full_n$N.log 2>&1; echo "dist_check resnet50 rc=$?"; grep -E "(OK|MISMATCH)|Error|error" gpurun_out/dist_check_ring_full_n$N.log | head
GQ_P2P=0 GQ_RING_PARTS=1 timeout 300 $RUN --master-port 29539 bench.py --gpus $N --mode ring --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ring_nccl_n$N.json 2> gpurun_out/bench_ring_nccl_n$N.err
timeout 300 $RUN --master-port 29534 bench.py --gpus $N --mode ring --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ring_p2p_n$N.json 2> gpurun_out/bench_ring_p2p_n$N.err
timeout 300 $RUN --master-port 29535 bench.py --gpus $N --mode ps --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ps_n$N.json 2> gpurun_out/bench_ps_n$N.err
python - <<PY
import json
for f in ("bench_ring_nccl", "bench_ring_p2p", "bench_ps"):
    try:
        d = json.loads(open("gpurun_out/%s_n$N.json" % f).read().strip().splitlines()[-1])
        print("%s N=%d: %.1f us/step  %.1f Gelem/s  enc %.1f dec %.1f e2e %.2f ms [%s]" % (f, d["n_gpus"], d["ms_per_step"]*1e3, d["value"]/1e9, d["roofline"]["encode_ms"]*1e3, d["roofline"]["decode_ms"]*1e3, d["e2e"]["ms_per_step"], d["config"]["exchange"][:70]))
    except Exception as e:
        print(f, "parse failed", e); print(open("gpurun_out/%s_n$N.err" % f).read()[-500:])
PY
