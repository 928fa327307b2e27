"""GPU: the multi-rank ps path (one user per rank, peer-to-peer exchange of packed records through
CUDA IPC + barrier kernel + gather kernel) exercised with TWO PROCESSES ON ONE GPU, so that it is
covered by a single-GPU `pytest -m gpu` run.  Process-group plumbing uses gloo (NCCL refuses two
ranks on one device); the records themselves never leave the GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, mode):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["GQ_P2P_MODE"] = mode
    os.environ["GQ_P2P_ALLOC"] = "ipc"     # two ranks on ONE device: CUDA IPC (symmetric memory / NVLS need one GPU per rank)
    import torch.distributed as dist
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gq_b200
    from oracle import gq_oracle as O
    from util import FCN_SHAPES, codebook, gen_input, make_args, torch_uniform_stream

    dev = torch.device("cuda", 0)
    shapes = FCN_SHAPES + [(64, 3, 3, 3)]
    sizes = [int(np.prod(s)) for s in shapes]
    a = make_args(mode="ps", num_users=world)
    params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
    ok = q.p2p is not None
    per_user = sum(n // 16 for n in sizes if n > 1000)
    codecs = [O.HSQ(n, s, codebook(16, 256), 6, True) if n > 1000 else O.Identity() for n, s in zip(sizes, shapes)]
    for it in range(3):     # three steps: both record parities and their reuse
        grads = [[gen_input(9000 + 100 * it + 10 * u + i, n).reshape(s) for i, (n, s) in enumerate(zip(sizes, shapes))]
                 for u in range(world)]
        stream = torch_uniform_stream(500 + it, per_user * world)
        for p, g in zip(params, grads[rank]):
            p.grad = torch.from_numpy(g).to(dev)
        parts, _ = q.plan.split_uniform_stream(stream[rank * per_user:(rank + 1) * per_user])
        q.record(rank, epoch=1, uniforms=parts)
        q.apply()
        torch.cuda.synchronize()
        ref = O.ps_step(codecs, grads, O.UniformStream(stream))
        for p, r in zip(params, ref):
            got = p.grad.data.cpu().numpy()
            ok = ok and np.array_equal(got, r.reshape(got.shape))
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as fh:
        fh.write("1" if ok else "0")
    dist.barrier()
    if q.p2p is not None:
        q.p2p.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["push", "gather", "direct"])
def test_ps_peer_to_peer_two_processes_one_gpu(tmp_path, mode):
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True,
                       start_method="spawn")
    for r in range(world):
        assert open(tmp_path / ("rank%d.txt" % r)).read() == "1", "rank %d: mismatch or P2P not active" % r
