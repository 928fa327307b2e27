"""Peer-to-peer exchange of packed records (ps topology, one user per GPU, one NVSwitch node).

Each rank owns one cudaMalloc'ed, IPC-exported buffer  [record 0 | record 1 | flags]  and maps
every peer's buffer.  A step writes the local record of parity p, runs the barrier kernel
(gq_peer_barrier) and then decodes all users' records of parity p straight out of peer memory.
Double buffering makes one barrier per step sufficient: a rank can only overwrite record p two
steps later, after every peer has passed the next barrier, i.e. finished reading it.
"""
import ctypes

import torch
import torch.distributed as dist

from .. import _lib


class _CudaBuffer:
    """Exposes a raw device pointer to torch (zero copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class PeerRecords:
    FLAG_BYTES = 256

    def __init__(self, record_bytes, rank, world, device):
        assert world <= 8
        self.record_bytes, self.rank, self.world, self.device = record_bytes, rank, world, device
        total = 2 * record_bytes + self.FLAG_BYTES
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_char * 64)()
        _lib.call("gq_ipc_alloc", total, ctypes.byref(ptr), ctypes.cast(handle, ctypes.c_void_p))
        self.local_ptr = ptr.value
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw))
        self.base = []
        self._opened = []
        for r in range(world):
            if r == rank:
                self.base.append(self.local_ptr)
            else:
                pp = ctypes.c_void_p()
                hbuf = (ctypes.c_char * 64).from_buffer_copy(handles[r])
                _lib.call("gq_ipc_open", ctypes.cast(hbuf, ctypes.c_void_p), ctypes.byref(pp))
                self.base.append(pp.value)
                self._opened.append(pp.value)
        self._holder = _CudaBuffer(self.local_ptr, 2 * record_bytes)
        # [2, record_bytes] uint8 view of the local double buffer: the plan encodes into row `parity`
        self.records = torch.as_tensor(self._holder, device=device).view(2, record_bytes)
        self._flag_ptrs = (ctypes.c_void_p * world)(*[b + 2 * record_bytes for b in self.base])
        self.epoch = 0
        self.step = 0
        dist.barrier()   # every rank has mapped every buffer before anyone writes flags

    @property
    def parity(self):
        return self.step & 1

    def barrier(self):
        self.epoch += 1
        _lib.call("gq_peer_barrier", ctypes.cast(self._flag_ptrs, ctypes.c_void_p), self.rank, self.world,
                  self.epoch, _lib.stream())

    def gather(self, dst):
        """Pull every user's record of the current parity into dst ([U, record_bytes], local)."""
        srcs = (ctypes.c_void_p * self.world)(*[b + self.parity * self.record_bytes for b in self.base])
        _lib.call("gq_peer_gather", dst.data_ptr(), ctypes.cast(srcs, ctypes.c_void_p), self.record_bytes,
                  dst.stride(0), self.world, _lib.stream())

    def user0_record_ptr(self):
        return self.base[0] + self.parity * self.record_bytes

    def user_offsets(self):
        """host int64 array: byte distance of user u's record from user 0's (any sign)."""
        return (ctypes.c_int64 * self.world)(*[b - self.base[0] for b in self.base])

    def advance(self):
        self.step += 1

    def close(self):
        try:
            torch.cuda.synchronize()
            for p in self._opened:
                _lib.call("gq_ipc_close", p)
            _lib.call("gq_ipc_free", self.local_ptr)
        except Exception:  # noqa: BLE001 - best effort at interpreter shutdown
            pass
        self._opened = []
