#!/usr/bin/env python
"""bench.py -- headline benchmark of the HSQ gradient-compression hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the 161 gradient tensors of the reference's CIFAR
ResNet-50 (23 520 842 elements; 76 tensors / 23 498 432 elements compressed, the rest
identity), HSQ d=16, K=256 learned codebook, 6-bit norms, one simulated user per GPU.
One step on every rank = fused encode of the local gradient into its packed record ->
NCCL all-gather of the records (N > 1) -> fused decode-and-average over the N records.
Metric = gradient elements per second through the whole job (N * elements / step time).

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks; inputs rotate over buffers larger than L2.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "HSQ gradient elems/s (encode+allgather+decode)"
UNIT = "elements/s"


def resnet50_shapes():
    from util import resnet50_shapes as f
    return f()


def make_args(num_users):
    from types import SimpleNamespace
    return SimpleNamespace(c_dim=16, k_bit=8, n_bit=6, no_cuda=False, random=True, cr=256, ef=False,
                           two_phase=False, mode="ps", scale="exp", num_users=num_users)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the search kernel from the committed
    `ncu --set full` capture (profiles/), per launch; None if the summary is missing."""
    import csv
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prof_tc_*_summary.csv"))):
        rd = wr = None
        with open(path) as fh:
            for row in csv.reader(fh):
                if len(row) >= 3 and row[0] == "dram__bytes_read.sum":
                    rd = float(row[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(row[1], 1.0)
                if len(row) >= 3 and row[0] == "dram__bytes_write.sum":
                    wr = float(row[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(row[1], 1.0)
        if rd is not None and wr is not None:
            best = (rd + wr, os.path.basename(path))
    return best


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------ clocks ---
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed loops run."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------- CPU baselines ---
def cpu_oracle_pass(shapes, n_users, sample_elems=None, repeats=3):
    """The oracle port (C + OpenMP) of the reference path on the host cores: per user,
    decompress(compress(g)) over every tensor, then the mean over users
    (ps_quantizer.py:27-65).  Returns (elements/s, seconds, elements, threads)."""
    from oracle import gq_oracle as O
    cb = O.normalize(O.fvecs_read(os.path.join(
        ROOT, "gradient-quantization_b200", "codebooks", "learned_codebook", "angular_dim_16_Ks_256.fvecs")))[1]
    sizes = [int(np.prod(s)) for s in shapes]
    picked, total = [], 0
    for s, n in zip(shapes, sizes):
        if sample_elems is not None and total + n > sample_elems and picked:
            continue
        picked.append((s, n))
        total += n
    rs = np.random.RandomState(1)
    grads = [[(rs.standard_normal(n) * 0.01).astype(np.float32).reshape(s) for s, n in picked]
             for _ in range(n_users)]
    codecs = [O.HSQ(n, s, cb, 6, True) if n > 1000 else O.Identity() for s, n in picked]
    n_draws = sum(n // 16 for _, n in picked if n > 1000) * n_users
    draws = rs.random_sample(n_draws).astype(np.float32)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.ps_step(codecs, grads, O.UniformStream(draws))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    elems = total * n_users
    return elems / best, best, elems, int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


def cpu_torch_ops_pass(sample_elems=2_000_000, repeats=2):
    """How the reference itself computes HSQ on a CPU: torch.mm -> abs -> argmax -> gather ->
    min/max affine stochastic rounding -> gather * norm (nearest_neighbor_compressor.py:63-90,
    probabilistic_scalar_compressor.py:12-33), restated with the same torch ops, all host threads."""
    import torch
    from oracle import gq_oracle as O
    cb = torch.from_numpy(O.normalize(O.fvecs_read(os.path.join(
        ROOT, "gradient-quantization_b200", "codebooks", "learned_codebook", "angular_dim_16_Ks_256.fvecs")))[1])
    torch.set_num_threads(os.cpu_count() or 1)
    n = sample_elems // 16 * 16
    g = torch.randn(n) * 0.01
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        v = g.view(-1, 16)
        p = torch.mm(cb, v.transpose(0, 1)).transpose(0, 1)
        codes = torch.argmax(torch.abs(p), dim=1)
        u = p.gather(dim=1, index=codes.view(-1, 1)).view(-1)
        lb, ub = torch.min(u), torch.max(u)
        scaled = torch.abs((u - lb) / (ub - lb)) * 64
        l = torch.clamp(scaled, 0, 63).type(torch.int32)
        l += ((scaled - l.float()) > torch.rand(l.size())).type(torch.int32)
        norms = l.float() * (ub - lb) / 64 + lb
        out = cb[codes.long()] * norms.view(-1, 1)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    del out
    return n / best, torch.get_num_threads()


def run_reference_arm(a):
    """--impl reference: the reference path on the host CPU (oracle port; the reference is
    pure Python/PyTorch, so there is no compiled oracle/_ref).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shapes = resnet50_shapes()
    n_users = a.gpus
    sample = 23_520_842 if n_users == 1 else max(23_520_842 // n_users, 2_000_000)
    for _ in range(max(a.warmup, 1) - 1):
        cpu_oracle_pass(shapes, n_users, sample, repeats=1)
    times, elems, threads = [], 0, 1
    for _ in range(max(min(a.steps, 5), 1)):
        v, dt, elems, threads = cpu_oracle_pass(shapes, n_users, sample, repeats=1)
        times.append(dt)
    t = float(np.mean(times))
    value = elems / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": len(times), "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-50 (reference CIFAR variant) gradient tensors, HSQ d=16 K=256 n=6, "
                               "%d simulated users on the host CPU, ps record/apply" % n_users,
                   "sample_elements_per_user": elems // n_users},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d users x %d elements per step (C/OpenMP oracle port of the reference path)"
                                   % (n_users, elems // n_users)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------- GPU arm ---
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--algo", type=str, default="auto", choices=["auto", "exact", "tc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist

    import gq_b200
    from gq_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and a.gpus > 1:
        raise SystemExit("bench.py --gpus %d needs WORLD_SIZE == %d (launch with torch.distributed.run)"
                         % (a.gpus, a.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(a.warmup, 3), max(a.steps, 1)

    shapes = resnet50_shapes()
    args = make_args(world)
    args.hsq_algo = {"auto": _lib.ALGO_AUTO, "exact": _lib.ALGO_EXACT, "tc": _lib.ALGO_TC}[a.algo]
    params = [torch.nn.Parameter(torch.empty(0, device=dev)) for _ in shapes]
    for p, s in zip(params, shapes):
        p.data = torch.zeros(s, device=dev)
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, args)
    plan = q.plan
    n_total = plan.total_elems()
    g_hsq = plan.groups[0]

    # synthetic gradients, resident in HBM, rotating so that the inputs exceed L2 (126 MB)
    ROT = 4
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)
    inputs = [torch.randn(plan.arena_elems, device=dev, generator=gen) * 0.01 for _ in range(ROT)]
    outputs = [torch.empty(plan.arena_elems, device=dev) for _ in range(ROT)]

    def step(i):
        q.encode_local(rank, src=inputs[i % ROT])              # fused encode into the local packed record
        q.exchange_and_decode(out=outputs[i % ROT])             # P2P barrier (or NCCL all-gather) + fused decode

    # my own kernels per exchange: push/gather + barrier kernel (peer-to-peer); the NCCL all-gather is not mine
    exchange_launches = 0
    if q.p2p is not None:
        exchange_launches = {"push": 2, "gather": 2, "direct": 1}[q.p2p_mode]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    for i in range(W):
        step(i)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / K
    value = world * n_total / (ms_per_step * 1e-3)

    # ---- per-kernel timing of the dominant kernel (HSQ search), same rotation ----
    st = _lib.stream()
    ws = plan.workspace
    codes_ptr = plan.records[0].data_ptr() + g_hsq.codes_off

    def search_only(i):
        _lib.call("gq_hsq_search", inputs[i % ROT].data_ptr() + g_hsq.arena_off * 4, g_hsq.n_chunks, g_hsq.dim,
                  g_hsq.codebook.data_ptr(), g_hsq.K, codes_ptr, g_hsq.code_bytes, plan.u_scratch.data_ptr(),
                  g_hsq.seg_start.data_ptr(), g_hsq.n_seg, None, ws.data_ptr() + 4096, ws.numel() - 4096,
                  args.hsq_algo, st)

    def decode_only(i):
        q.decode_exchanged(outputs[i % ROT])

    def time_loop(fn, iters):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(iters):
            fn(i)
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / iters

    search_ms = time_loop(search_only, K)
    decode_ms = time_loop(decode_only, K)
    encode_ms = time_loop(lambda i: plan.encode(0, src=inputs[i % ROT]), K)

    # ---- end to end through the public quantizer API with HOST buffers ----
    # Every step copies its gradient from pinned host memory (H2D) and reads the averaged gradient
    # back to pinned host memory (D2H).  The three stages are software-pipelined over two device
    # buffers: H2D of step i+1 and D2H of step i-1 overlap the codec work of step i (PCIe is full
    # duplex), which is how a training loop would feed it.
    host_in = [torch.empty(plan.arena_elems, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h in host_in:
        h.copy_(inputs[0].cpu())
    host_out = [torch.empty(plan.arena_elems, dtype=torch.float32).pin_memory() for _ in range(2)]
    dev_in = [torch.empty(plan.arena_elems, device=dev) for _ in range(2)]
    dev_out = [torch.empty(plan.arena_elems, device=dev) for _ in range(2)]
    s_h2d, s_d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_in_free = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    ev_out_free = [torch.cuda.Event() for _ in range(2)]

    def issue_h2d(i):
        b = i % 2
        with torch.cuda.stream(s_h2d):
            s_h2d.wait_event(ev_in_free[b])
            dev_in[b].copy_(host_in[b], non_blocking=True)
            ev_in[b].record(s_h2d)

    def e2e_step(i):
        b = i % 2
        main.wait_event(ev_in[b])
        for p, v in zip(params, plan.views(dev_in[b])):
            p.grad = v                                            # this step's gradient (device views)
        q.record(rank, epoch=1)                                   # public API: gather + fused encode
        ev_in_free[b].record(main)
        q.apply()                                                 # all-gather + decode-and-average
        main.wait_event(ev_out_free[b])
        dev_out[b].copy_(plan.arena, non_blocking=True)
        ev_out[b].record(main)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_out[b])
            host_out[b].copy_(dev_out[b], non_blocking=True)      # D2H of the averaged gradient
            ev_out_free[b].record(s_d2h)

    for b in range(2):
        ev_in_free[b].record(main)
        ev_out_free[b].record(main)
    KE = max(min(K, 20), 4)

    def e2e_loop(n):
        issue_h2d(0)
        for i in range(n):
            if i + 1 < n:
                issue_h2d(i + 1)
            e2e_step(i)
        main.wait_stream(s_d2h)

    e2e_loop(3)
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    e2e_loop(KE)
    b1.record()
    barrier()
    e2e_ms = b0.elapsed_time(b1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_ms /= KE
    e2e_value = world * n_total / (e2e_ms * 1e-3)
    for p, v in zip(params, plan.views()):
        p.grad = v

    # keep the same loop running (~0.6 s) so that the sampler sees the clocks under load; the number
    # of extra steps is derived from the all-reduced step time, hence identical on every rank
    # (the steps contain collectives / peer barriers)
    n_extra = int(min(max(0.6 / (ms_per_step * 1e-3), 50), 20000))
    for i in range(n_extra):
        step(i)
        if i % 200 == 199:
            torch.cuda.synchronize()
    barrier()
    clocks = sampler.finish()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    traffic = ncu_traffic_bytes()
    hbm_peak = float(peaks["hbm_gbs"])
    # algorithmic bytes of the search kernel: read 4 B/elem, write 1 B code + 4 B u per chunk
    # (the 1-byte norm code is written by the quantize kernel from u; BASELINE.md counts
    #  4 + 2/d for the fused encode; both are reported)
    alg_bytes_search = g_hsq.n * 4 + g_hsq.n_chunks * (1 + 4)
    achieved = alg_bytes_search / (search_ms * 1e-3) / 1e9
    alg_bytes_step = n_total * (8 + 2.0 * (world + 1) / 16)
    roofline = {
        "bound": "hbm", "kernel": "hsq_search (%s)" % a.algo, "achieved": achieved, "peak": hbm_peak,
        "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic[0] if traffic else None,
        "traffic_source": traffic[1] if traffic else None, "peak_source": peak_kind,
        "kernel_ms": search_ms, "algorithmic_bytes_per_launch": alg_bytes_search,
        "tensor_flops_per_launch": 2.0 * g_hsq.K * g_hsq.n,
        "tensor_tflops_achieved": 2.0 * g_hsq.K * g_hsq.n / (search_ms * 1e-3) / 1e12,
        "step_hbm_frac": alg_bytes_step / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
        "encode_ms": encode_ms, "decode_ms": decode_ms,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-50 (reference CIFAR variant) 161 gradient tensors, 23520842 elements/user, "
                               "HSQ d=16 K=256 n=6, one user per GPU, ps encode+allgather+decode-mean",
                   "elements_per_user": n_total, "compressed_elements": plan.compressed_elems(),
                   "users": world, "wire_bytes_per_user": plan.wire_bytes(), "algo": a.algo,
                   "exchange": ("none (1 user)" if world == 1 else
                                ("peer-to-peer (%s%s): packed records cross NVLink through peer-mapped memory, barrier kernel"
                                 % (q.p2p_mode, ", NVLS multicast stores" if (q.p2p_mode == "push" and q.p2p.mc_base) else "")
                                 if q.p2p is not None else "NCCL all-gather of packed records")),
                   "l2": "inputs/outputs rotate over %d buffers of %.0f MB each (> 126 MB L2)"
                         % (ROT, plan.arena_elems * 4 / 1e6)},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": plan.arena_elems * 4, "d2h_bytes_per_step": plan.arena_elems * 4,
                "api": "PSQuantizer.record(rank)/apply() on gradients copied from pinned host memory each step, "
                       "averaged gradient copied back to pinned host memory each step",
                "pipelining": "double-buffered: H2D of step i+1 and D2H of step i-1 overlap step i"},
        "gpu_launches": K * (plan.launches_per_encode() + plan.launches_per_decode(world) + exchange_launches),
        "clocks": clocks,
    }
    if not a.no_cpu_baseline and world == 1:   # reported at N = 1 only (torchrun pins OMP to one thread per rank)
        v, dt, elems, threads = cpu_oracle_pass(shapes, 1, None, repeats=3)
        tv, tthreads = cpu_torch_ops_pass()
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "1 user-pass (encode+decode) of the full 23520842-element gradient, best of 3, "
                      "C/OpenMP oracle port (%.3f s)" % dt,
            "torch_ops_value": tv, "torch_ops_threads": tthreads,
            "torch_ops_sample": "2.0M elements, reference's torch.mm/abs/argmax/gather op sequence on CPU",
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
