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
                 max_persons: int = 32, max_peaks: int = _abi.MAX_PEAKS, strict: bool = True):
        from ._cuda_backend import CudaBackend          # raises without CUDA / the library
        self.backend = CudaBackend()
        self.model = model
        self.camera = camera
        self.config = config or DecodeConfig()
        self.input_size = input_size
        # the network's third head has num_limbs + 1 planes; joint j reads plane j (...mpreal_ablation.py:212-215)
        self.params = _abi.make_decode_params(self.config, camera, input_size=input_size, max_peaks=max_peaks,
                                              max_persons=max_persons, depth_channels=model.num_limbs + 1)
        #: the reference's lists are unbounded; max_peaks / max_persons are device capacities.  strict: collect() raises
        #: OverflowError when a frame hit one of them (its poses would differ from the reference's); strict=False
        #: leaves the check of out["flags"] to the caller.
        self.strict = strict
        self._out = None
        self._x_dev = None
        self._slots = None
        self.inject = None          # optional (heat, paf, depth) device tensors decoded INSTEAD of the network's maps

    # ---------------------------------------------------------------------------------------
    NSLOT = 3      # batches in flight: H2D of batch i+2, forward of batch i+1 and decode + D2H of batch i overlap

    def _buffers(self, B):
        if self._out is None or self._out["n_person"].shape[0] != B:
            from ._cuda_backend import alloc_decode_out
            self._slots = []
            for _ in range(self.NSLOT):
                out = alloc_decode_out(B, self.params)
                self._slots.append({
                    "out": out,
                    "x": torch.empty((B, 1, self.input_size, self.input_size), dtype=torch.float32, device="cuda"),
                    "host": torch.empty(out["_records"].shape, dtype=torch.uint8).pin_memory(),
                    "h2d": torch.cuda.Event(), "done": torch.cuda.Event(), "busy": False,
                })
            self._out = self._slots[0]["out"]
            self._x_dev = self._slots[0]["x"]
            self._copy_stream = torch.cuda.Stream()
            self.decode_stream = torch.cuda.Stream()
            self._next = 0
        return self._out

    def infer_device(self, x_dev, out=None, after=None, _evs=None):
        """x_dev [B,1,H,W] fp32 CUDA -> dict of device record tensors (no synchronisation).

        The forward runs on the current stream; decode + lift run on a second stream that waits on the forward's
        completion event, so the (latency-bound, low-occupancy) decode of batch i overlaps the forward of batch i+1.
        Work later enqueued on ``self.decode_stream`` (D2H, all-gather) is ordered after the decode; callers that read
        the records from another stream must wait on ``out["_ready"]``.  ``after``: optional callable run on the decode
        stream right after the decode (used for the D2H copy / the collective)."""
        B = x_dev.shape[0]
        self._buffers(B)
        out = self._out if out is None else out
        main = torch.cuda.current_stream()
        if _evs is not None:
            _evs[0].record(main)            # bench hook: start of the forward on its stream
        (paf, heat, depth), _ = self.model(x_dev)
        if _evs is not None:
            _evs[1].record(main)            # bench hook: end of the forward
        fwd_done = torch.cuda.Event()
        fwd_done.record(main)
        if self.inject is not None:
            heat, paf, depth = self.inject
        ds = self.decode_stream
        ds.wait_event(fwd_done)
        with torch.cuda.stream(ds):
            if _evs is not None:
                _evs[2].record(ds)          # bench hook: start of the decode on the decode stream
            self.backend.decode_device(heat, paf, depth, self.params, out)
            for t in (heat, paf, depth):
                t.record_stream(ds)
            res = after(out) if after is not None else None
            ready = torch.cuda.Event()
            ready.record(ds)
        out["_ready"] = ready
        out["_after"] = res
        return out

    def submit(self, frames, after=None):
        """Asynchronous half of ``infer``: enqueue H2D (copy stream) -> forward -> decode -> D2H for one batch of HOST
        frames and return a ticket.  Up to NSLOT batches may be in flight; ``collect`` them in submission order.
        ``after(out)``: optional callable enqueued on the decode stream behind the D2H copy (e.g. ``gather_records``)."""
        x = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames, np.float32))
        B = x.shape[0]
        self._buffers(B)
        slot = self._slots[self._next % self.NSLOT]
        if slot["busy"]:
            raise RuntimeError("collect() the oldest batch before submitting a %dth one" % (self.NSLOT + 1))
        self._copy_stream.wait_event(slot["done"])          # the previous user of this slot has finished with x / host
        with torch.cuda.stream(self._copy_stream):
            slot["x"].copy_(x, non_blocking=True)
            slot["h2d"].record(self._copy_stream)
        torch.cuda.current_stream().wait_event(slot["h2d"])
        # one D2H transfer for all record fields, on the decode stream right behind the decode
        def tail(o):
            slot["host"].copy_(o["_records"], non_blocking=True)
            return after(o) if after is not None else None
        out = self.infer_device(slot["x"], slot["out"], after=tail)
        slot["done"] = out["_ready"]
        slot["busy"] = True
        self._next += 1
        return (slot, B)

    def collect(self, ticket):
        """Wait for a submitted batch; returns NumPy views of its pose records (valid until the slot is reused)."""
        slot, B = ticket
        slot["done"].synchronize()
        slot["busy"] = False
        rec = unpack_records(slot["host"], slot["out"]["_layout"], B)
        if self.strict and rec["flags"].any():
            from .decode import _raise_on_overflow
            _raise_on_overflow(rec["flags"])
        return rec

    def infer(self, frames):
        """The user-facing call: ``frames`` [B,1,H,W] fp32 on the HOST (NumPy array or, to avoid a staging copy,
        a pinned torch tensor) -> dict of NumPy pose records.  Includes H2D of the frames and D2H of the records."""
        return self.collect(self.submit(frames))

    def h2d_bytes(self, B):
        return B * self.input_size * self.input_size * 4

    def d2h_bytes(self, B):
        self._buffers(B)
        return int(self._out["_records"].numel())


# ---------------------------------------------------------------------------------------------
# the collective
# ---------------------------------------------------------------------------------------------
def unpack_records(buf, layout, B, world=None):
    """Views of the record fields inside a packed byte buffer (NumPy for host buffers, torch for device buffers).
    With ``world`` the buffer holds `world` rank chunks back to back and every view gets a leading rank axis."""
    is_torch = isinstance(buf, torch.Tensor) and buf.is_cuda
    res = {}
    if world is None:
        base = buf if is_torch else (buf.numpy() if isinstance(buf, torch.Tensor) else buf)
        for name, dt, shp, off, nb in layout:
            if is_torch:
                res[name] = base[off:off + nb].view(dt).reshape((B,) + tuple(shp))
            else:
                res[name] = base[off:off + nb].view(_NP[dt]).reshape((B,) + tuple(shp))
        return res
    chunk = buf.shape[0] // world
    base = buf.reshape(world, chunk) if is_torch else (buf.numpy() if isinstance(buf, torch.Tensor) else buf).reshape(world, chunk)
    for name, dt, shp, off, nb in layout:
        if is_torch:
            res[name] = base[:, off:off + nb].contiguous().view(dt).reshape((world * B,) + tuple(shp))
        else:
            res[name] = np.ascontiguousarray(base[:, off:off + nb]).view(_NP[dt]).reshape((world * B,) + tuple(shp))
    return res


_NP = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32, torch.int16: np.int16}


def gather_records(out, group=None, unpack=True):
    """All-gather the pose records of equal-sized batch shards: ONE collective on the packed record bytes
    (``out["_records"]``, see _cuda_backend.alloc_decode_out); rank r's frames land at [r*B, (r+1)*B).
    Works on CUDA buffers (NCCL over NVLink) and CPU buffers (gloo, used by the tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rec = out["_records"]
    rec = rec if isinstance(rec, torch.Tensor) else torch.from_numpy(rec)
    full = out.get("_gathered")
    if full is None or full.shape[0] != world * rec.shape[0] or full.device != rec.device:
        # one buffer per output slot, reused every step (all ranks must hold equal-sized shards: see shard())
        full = out["_gathered"] = torch.empty((world * rec.shape[0],), dtype=torch.uint8, device=rec.device)
    dist.all_gather_into_tensor(full, rec, group=group)
    if not unpack:
        return full
    B = out["n_person"].shape[0]
    return unpack_records(full, out["_layout"], B, world)


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
    """Contiguous batch slice of rank `rank`.  Every rank gets ceil(n / world) slots; ranks past the
    end get a shorter (possibly empty) slice -- pad it with ``pad_shard`` before a gather, which needs equal sizes."""
    per = (n_items + world - 1) // world
    return slice(min(rank * per, n_items), min((rank + 1) * per, n_items))


def pad_shard(x, n_items: int, world: int):
    """Pad a rank's shard (NumPy array or tensor, frames on axis 0) with zero frames up to ceil(n / world) so that all
    ranks run the same batch size and ``gather_records`` sees equal-sized buffers.  Returns (padded, n_valid).  A zero
    depth frame decodes to no persons; callers drop the padding with ``unpad_gathered``."""
    per = (n_items + world - 1) // world
    n = x.shape[0]
    if n == per:
        return x, n
    if isinstance(x, torch.Tensor):
        pad = torch.zeros((per - n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        return torch.cat([x, pad], 0), n
    return np.concatenate([x, np.zeros((per - n,) + x.shape[1:], x.dtype)], 0), n


def unpad_gathered(rec: dict, n_items: int):
    """Drop the padding frames of a gathered record dict (rank r's frames sit at [r * per, r * per + n_r))."""
    return {k: v[:n_items] for k, v in rec.items()}
