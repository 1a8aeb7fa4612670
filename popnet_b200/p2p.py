"""The multi-GPU record exchange: peer-mapped gather buffers over NVLink (host half).

One process per GPU (torch.distributed is the rendezvous: handle exchange and barriers only).  Every rank owns one
peer-visible device block (``popnet_p2p_alloc``: cudaMalloc + CUDA IPC handle) laid out, per pipeline slot, as

    gather[world][records_bytes]    rank r's pose records of the step -- identical on every rank after the step
    arrive[world] (uint64)          step tags; arrive[r] is written by rank r's assembly kernel

and maps the blocks of all peers.  ``PoseEstimator`` points the record fields of its decode output INTO
``gather[rank]`` of its own block; ``popnet_decode_push`` makes the assembly kernel store every record value into the
same place of every peer's block too and publish a tag; ``popnet_p2p_wait`` (one warp) returns when all tags of the
step are in.  No all-gather kernel exists: the transfer rides on the stores of the kernel that produces the data, and
no SMs are held while waiting for the slowest rank (include/popnet_b200.h, PopnetPeerPush).

The reference has no multi-GPU path to mirror (DataParallel on one device, ...mpreal_ablation.py:140-141).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi, _lib


class _Blob:
    """Wrap a raw device pointer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerGather:
    def __init__(self, records_bytes: int, nslot: int, group=None, timeout_ms: int = 2000):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.PopnetError("PeerGather needs an initialised torch.distributed process group")
        self.lib = _lib.get()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > _abi.MAX_PEERS:
            raise _lib.PopnetError("at most %d GPUs of one NVSwitch box (got world size %d)" % (_abi.MAX_PEERS, self.world))
        self.records_bytes = (int(records_bytes) + 255) // 256 * 256
        self.nslot = nslot
        self.timeout_ms = timeout_ms
        self.slot_bytes = self.world * self.records_bytes + 256          # gather chunks + the arrive array (8 x u64, padded)
        self.bytes = nslot * self.slot_bytes
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.check(self.lib.popnet_p2p_alloc(self.bytes, C.byref(ptr), handle), "popnet_p2p_alloc")
        self._own = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.base = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.base.append(self._own)
                continue
            p = C.c_void_p()
            _lib.check(self.lib.popnet_p2p_open(h, C.byref(p)), "popnet_p2p_open (rank %d)" % r)
            self.base.append(p.value)
        self.block = torch.as_tensor(_Blob(self._own, self.bytes), device="cuda")          # own block as a uint8 tensor
        # local bookkeeping words per slot: step counter (u64), CTA counter (u32), status (u32)
        self.state = torch.zeros((nslot, 4), dtype=torch.int64, device="cuda")
        self._push = [self._make_push(i) for i in range(nslot)]
        self.barrier()

    # ---- layout
    def _gather_ptr(self, r, slot):
        return self.base[r] + slot * self.slot_bytes

    def _arrive_ptr(self, r, slot):
        return self.base[r] + slot * self.slot_bytes + self.world * self.records_bytes

    def _make_push(self, slot):
        p = _abi.PeerPush()
        p.world, p.rank = self.world, self.rank
        for r in range(self.world):
            p.gather_base[r] = self._gather_ptr(r, slot)
            p.arrive[r] = self._arrive_ptr(r, slot)
        p.records_bytes = self.records_bytes
        st = self.state[slot].data_ptr()
        p.step, p.done_counter, p.status = st, st + 8, st + 16
        return p

    def push_args(self, slot):
        return self._push[slot]

    def local_records(self, slot, B=None, params=None):
        """uint8 view of this rank's chunk of its own gather buffer: the decode writes its record block here."""
        o = slot * self.slot_bytes + self.rank * self.records_bytes
        return self.block[o:o + self.records_bytes]

    def gathered(self, slot):
        """uint8 view of the slot's whole gather buffer: world chunks of records_bytes each, rank-major."""
        o = slot * self.slot_bytes
        return self.block[o:o + self.world * self.records_bytes]

    # ---- per step
    def wait_arrivals(self, slot):
        st = self.state[slot].data_ptr()
        _lib.check(self.lib.popnet_p2p_wait(C.c_void_p(self._arrive_ptr(self.rank, slot)), self.world, C.c_void_p(st),
                                            C.c_void_p(st + 16), self.timeout_ms,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "popnet_p2p_wait")

    def timed_out(self) -> bool:
        """True if any wait gave up on a peer (synchronises)."""
        return bool(self.state[:, 2].ne(0).any().item())

    def barrier(self):
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def close(self):
        if self._own is None:
            return
        self.barrier()
        for r, p in enumerate(self.base):
            if r != self.rank:
                self.lib.popnet_p2p_close(C.c_void_p(p))
        self.block = None
        self.lib.popnet_p2p_free(C.c_void_p(self._own))
        self._own = None
