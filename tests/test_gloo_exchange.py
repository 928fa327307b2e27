"""CPU, world_size 2, gloo: the exchange plumbing of the ps and ring topologies moves the
packed records exactly as the single-process simulation lays them out."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gq_b200
    from gq_b200.quantizers import exchange as xch
    from gq_b200.quantizers.fused import FusedPlan
    from util import FCN_SHAPES, make_args

    plan = FusedPlan(gq_b200.NearestNeighborCompressor, FCN_SHAPES, make_args(num_users=world),
                     torch.device("cpu"), world)
    rb = plan.record_bytes
    # every rank "encodes" a recognisable record into its own slot only
    plan.records.zero_()
    plan.records[rank] = torch.from_numpy(((np.arange(rb) * (rank + 3)) % 251).astype(np.uint8))
    xch.ps_all_gather(plan.records, rank)
    ok_ps = all(np.array_equal(plan.records[r].numpy(), ((np.arange(rb) * (r + 3)) % 251).astype(np.uint8))
                for r in range(world))

    # ring: rank r receives r-1's record, writes its own, forwards; last one is broadcast
    plan.records.zero_()
    xch.ring_receive_previous(plan.records, rank)
    prev_ok = True
    if rank > 0:
        prev_ok = np.array_equal(plan.records[rank - 1].numpy(), np.full(rb, 10 + rank - 1, np.uint8))
    plan.records[rank] = torch.full((rb,), 10 + rank, dtype=torch.uint8)
    xch.ring_send_next(plan.records, rank, world)
    xch.ring_broadcast_last(plan.records, world)
    last_ok = np.array_equal(plan.records[world - 1].numpy(), np.full(rb, 10 + world - 1, np.uint8))
    # pipelined ring: the plan cut into stages; every stage's byte ranges travel as one batched group and
    # together the stages cover every useful byte of the record exactly once
    n_parts = plan.make_parts(3)
    plan.records.zero_()
    covered = np.zeros(rb, dtype=np.int32)
    stage_ok = n_parts >= 2
    for p in range(n_parts):
        ranges = plan.part_byte_ranges(p)
        for a, b in ranges:
            covered[a:b] += 1
        xch.ring_receive_previous_part(plan.records, rank, ranges)
        if rank > 0:
            for a, b in ranges:
                stage_ok = stage_ok and bool((plan.records[rank - 1][a:b] == 20 + p).all())
        for a, b in ranges:
            plan.records[rank][a:b] = 20 + p
        xch.ring_send_next_part(plan.records, rank, world, ranges)
    stage_ok = stage_ok and covered.max() == 1 and int((covered == 1).sum()) >= plan.wire_bytes()
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as fh:
        fh.write("%d %d %d %d" % (ok_ps, prev_ok, last_ok, stage_ok))
    dist.destroy_process_group()


def test_ps_and_ring_exchange_gloo_world2(tmp_path):
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    for r in range(world):
        flags = open(tmp_path / ("rank%d.txt" % r)).read().split()
        assert flags == ["1", "1", "1", "1"], (r, flags)


def test_distributed_quantizer_requires_one_user_per_rank():
    """host-side validation only (no process group): world size is read lazily."""
    from gq_b200.quantizers._shared import dist_world
    assert dist_world() == (0, 1)
