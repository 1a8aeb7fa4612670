"""End-to-end depth-pose inference: frames -> rtpose_light3d forward -> decode -> 3D lift, batched on the
device, plus the one collective of the multi-GPU path.

This is the hot loop of the reference's eval scripts
(third_party_methods/evaluate/evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:161-316) with the
per-frame Python body replaced by three kernel launches per batch; the host<->device boundary is crossed
twice per batch, like the reference (`img.cuda()` :171 and the output copy :176-178), but what comes
back is the fixed-size pose records instead of the 185 KB/frame of raw maps.

Multi-GPU (one process per GPU, torch.distributed): frames are sharded by contiguous batch slices;
nothing is exchanged during forward/decode.  ``gather_records`` all-gathers the per-frame pose records
(NCCL over NVLink on GPUs, gloo on CPU tensors in the tests); evaluator counters are integers and are
summed with ``reduce_counts``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _abi
from .topology import MP3DHP, Camera, DecodeConfig

RECORD_KEYS = ("n_person", "flags", "person_peak", "person_score", "person_njoint", "pose2d", "pose3d", "pose_conf")


class PoseEstimator:
    def __init__(self, model, camera: Camera = MP3DHP, config: DecodeConfig | None = None, *, input_size: int = 224,
                 max_persons: int = 32, max_peaks: int = _abi.MAX_PEAKS):
        from ._cuda_backend import CudaBackend          # raises without CUDA / the library
        self.backend = CudaBackend()
        self.model = model
        self.camera = camera
        self.config = config or DecodeConfig()
        self.input_size = input_size
        self.params = _abi.make_decode_params(self.config, camera, input_size=input_size, max_peaks=max_peaks,
                                              max_persons=max_persons)
        self._out = None
        self._x_dev = None
        self._host = None
        self.inject = None          # optional (heat, paf, depth) device tensors decoded INSTEAD of the network's maps

    # ---------------------------------------------------------------------------------------
    def _buffers(self, B):
        if self._out is None or self._out["n_person"].shape[0] != B:
            from ._cuda_backend import alloc_decode_out
            self._out = alloc_decode_out(B, self.params)
            self._x_dev = torch.empty((B, 1, self.input_size, self.input_size), dtype=torch.float32, device="cuda")
            self._host = {k: torch.empty(self._out[k].shape, dtype=self._out[k].dtype).pin_memory() for k in RECORD_KEYS}
        return self._out

    def infer_device(self, x_dev):
        """x_dev [B,1,H,W] fp32 CUDA -> dict of device record tensors (no synchronisation)."""
        B = x_dev.shape[0]
        out = self._buffers(B)
        (paf, heat, depth), _ = self.model(x_dev)
        if self.inject is not None:
            heat, paf, depth = self.inject
        self.backend.decode_device(heat, paf, depth, self.params, out)
        return out

    def infer(self, frames):
        """The user-facing call: ``frames`` [B,1,H,W] fp32 on the HOST (NumPy array or, to avoid a staging copy,
        a pinned torch tensor) -> dict of NumPy pose records.  Includes H2D of the frames and D2H of the records."""
        x = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames, np.float32))
        B = x.shape[0]
        self._buffers(B)
        self._x_dev.copy_(x, non_blocking=True)
        out = self.infer_device(self._x_dev)
        for k in RECORD_KEYS:
            self._host[k].copy_(out[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        res = {k: self._host[k].numpy() for k in RECORD_KEYS}
        return res

    def h2d_bytes(self, B):
        return B * self.input_size * self.input_size * 4

    def d2h_bytes(self, B):
        self._buffers(B)
        return int(sum(self._out[k].numel() * self._out[k].element_size() for k in RECORD_KEYS))


# ---------------------------------------------------------------------------------------------
# the collective
# ---------------------------------------------------------------------------------------------
def gather_records(records, group=None):
    """All-gather fixed-size pose records of equal-sized batch shards; rank r's frames land at
    [r*B, (r+1)*B).  Works on CUDA tensors (NCCL) and CPU tensors (gloo).  The eight record arrays of a
    frame are packed into one byte row so that the step issues ONE collective (int16 fields are not a
    collective dtype in either backend anyway).  Returns a dict of tensors."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts, meta = [], []
    B = None
    for k in RECORD_KEYS:
        t = records[k]
        t = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t))
        if t.dtype == torch.uint32:
            t = t.view(torch.int32)
        t = t.contiguous()
        B = t.shape[0]
        row = t.reshape(B, -1).view(torch.uint8)
        parts.append(row)
        meta.append((k, t.dtype, tuple(t.shape[1:]), row.shape[1]))
    packed = torch.cat(parts, dim=1).contiguous()
    full = torch.empty((world * B, packed.shape[1]), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(full, packed, group=group)
    out, off = {}, 0
    for k, dt, shape, nbytes in meta:
        out[k] = full[:, off:off + nbytes].contiguous().view(dt).reshape((world * B,) + shape)
        off += nbytes
    return out


def reduce_counts(counts, group=None):
    """Sum integer evaluator counters (hit_cnt, valid_cnt, n_gt, n_pos, samples) over ranks."""
    import torch.distributed as dist
    out = {}
    for k, v in counts.items():
        t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, np.int64))
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out[k] = t
    return out


def shard(n_items: int, rank: int, world: int):
    """Contiguous batch slice of rank `rank`."""
    per = (n_items + world - 1) // world
    return slice(min(rank * per, n_items), min((rank + 1) * per, n_items))
