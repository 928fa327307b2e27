"""The exchange step of the two topologies, on packed records only.

records is the [U, record_bytes] uint8 buffer of a FusedPlan; with one user per rank,
rank r's own record sits in row r.  These helpers are device-agnostic (NCCL on GPUs,
gloo on CPU in the tests): they only move bytes."""
import torch.distributed as dist


def ps_all_gather(records, rank):
    """ps: every rank ends up with every user's packed record (one all-gather, in place:
    row `rank` is the send buffer).  Each rank then decodes-and-averages locally, in user
    order, so all ranks hold the identical averaged gradient without a second collective."""
    dist.all_gather_into_tensor(records.view(-1), records[rank])


def ring_receive_previous(records, rank):
    """ring: the running sum's packed record arrives from rank - 1 (ring_quantizer.py:31-32)."""
    if rank > 0:
        dist.recv(records[rank - 1], src=rank - 1)


def ring_send_next(records, rank, world):
    if rank + 1 < world:
        dist.send(records[rank], dst=rank + 1)


def ring_broadcast_last(records, world):
    """ring: the last hop's record is what every rank decodes (ring_quantizer.py:45-49)."""
    dist.broadcast(records[world - 1], src=world - 1)


def _part_ops(op, record, ranges, peer):
    return [dist.P2POp(op, record[a:b], peer) for a, b in ranges if b > a]


def ring_receive_previous_part(records, rank, ranges):
    """Pipelined ring: one stage's byte ranges of the previous hop's record, as ONE batched group of
    receives (the sections of a stage are not contiguous in the record)."""
    if rank > 0:
        for w in dist.batch_isend_irecv(_part_ops(dist.irecv, records[rank - 1], ranges, rank - 1)):
            w.wait()


def ring_send_next_part(records, rank, world, ranges):
    if rank + 1 < world:
        for w in dist.batch_isend_irecv(_part_ops(dist.isend, records[rank], ranges, rank + 1)):
            w.wait()
