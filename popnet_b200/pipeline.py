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
    """frames -> poses.  Each batch in flight owns a *slot*: the device input buffer, the six network output maps, the
    decode record buffer and its pinned host mirror.  The launch sequence of a slot -- forward (39 convolutions on three
    streams), then decode + lift (+ the record push to the peer GPUs) -- is captured ONCE into two CUDA graphs and
    replayed every step: one graph launch on the main stream for the forward, one on the decode stream for the decode,
    so the (latency-bound) decode of batch i runs under the forward of batch i+1 and the host issues two launches per
    step instead of ~45.  ``use_graphs=False`` issues the same launches eagerly (tests compare both)."""

    NSLOT = 3      # batches in flight: H2D of batch i+2, forward of batch i+1 and decode + D2H of batch i overlap

    def __init__(self, model, camera: Camera = MP3DHP, config: DecodeConfig | None = None, *, input_size=224,
                 max_persons: int = 32, max_peaks: int = _abi.MAX_PEAKS, strict: bool = True, use_graphs: bool = True,
                 peers=None, decode_priority: int = 0, reserve_sms: int = 8, decode_ctas: int | None = None):
        from ._cuda_backend import CudaBackend          # raises without CUDA / the library
        self.backend = CudaBackend()
        self.model = model
        self.camera = camera
        self.config = config or DecodeConfig()
        self.input_hw = (input_size, input_size) if isinstance(input_size, int) else tuple(input_size)
        self.input_size = self.input_hw[0]
        ds = self.config.downsample
        # the network's third head has num_limbs + 1 planes; joint j reads plane j (...mpreal_ablation.py:212-215)
        self.params = _abi.make_decode_params(self.config, camera, input_size=self.input_hw[1], max_peaks=max_peaks,
                                              max_persons=max_persons, depth_channels=model.num_limbs + 1,
                                              max_ctas=reserve_sms if decode_ctas is None else decode_ctas,
                                              grid_hw=(self.input_hw[0] // ds, self.input_hw[1] // ds))
        #: the reference's lists are unbounded; max_peaks / max_persons are device capacities.  strict: collect() raises
        #: OverflowError when a frame hit one of them (its poses would differ from the reference's); strict=False
        #: leaves the check of out["flags"] to the caller.
        self.strict = strict
        self.use_graphs = use_graphs
        #: the decode of batch i runs under the forward of batch i + 1, and a convolution CTA fills an SM's register file: a
        #: decode CTA on an SM keeps the next layer's CTA off it and the whole layer waits.  The persistent conv grids
        #: therefore leave `reserve_sms` SMs (a multiple of 4, 0..28) to the decode (POPNET_TUNE_RESERVE_SMS; measured, same
        #: box: 1.060 vs 1.084 ms per step with 8, forward alone 1.016 vs 1.009 ms).  Only set if the model's tuning
        #: word does not choose a reserve itself.
        if reserve_sms and not (model.tuning & _abi.TUNE_RESERVE_SMS(7)):
            model.tuning |= _abi.TUNE_RESERVE_SMS(reserve_sms // 4)
        self.reserve_sms = 4 * ((model.tuning >> 9) & 7)
        self.decode_priority = decode_priority      # CUDA stream priority of the decode stream (0 = default, -1 = high)
        self.peers = peers          # optional p2p.PeerGather: the multi-GPU record exchange, fused into the decode
        self._slots = None
        self._B = None
        self.inject = None          # optional (heat, paf, depth) device tensors decoded INSTEAD of the network's maps
                                    # (decode-only tests; eager mode only)

    # ---------------------------------------------------------------------------------------
    def _buffers(self, B):
        if self._slots is not None and self._B == B:
            return self._slots
        from ._cuda_backend import alloc_decode_out
        H, W = self.input_hw
        self._prepared = self.model.prepare(B, H, W)
        self._slots = []
        for i in range(self.NSLOT):
            out = alloc_decode_out(B, self.params, records=None if self.peers is None else self.peers.local_records(i, B, self.params))
            self._slots.append({
                "index": i, "out": out, "x": torch.zeros((B, 1, H, W), dtype=torch.float32, device="cuda"),
                "maps": self.model.alloc_maps(B, H, W),
                "host": torch.empty(out["_records"].shape, dtype=torch.uint8).pin_memory(),
                "h2d": torch.cuda.Event(), "fwd_done": torch.cuda.Event(), "done": torch.cuda.Event(), "busy": False,
                "graphs": None,
            })
        self._B = B
        self._copy_stream = torch.cuda.Stream()
        self.decode_stream = torch.cuda.Stream(priority=self.decode_priority)
        self._capture_stream = torch.cuda.Stream()
        # (kernel nodes of a graph keep the priority of the stream they were captured on)
        self._capture_stream_dec = torch.cuda.Stream(priority=self.decode_priority) if self.decode_priority else self._capture_stream
        self._next = 0
        return self._slots

    def refresh(self):
        """Drop the captured graphs and buffers (after editing the model's weights or changing operand_dtype)."""
        self._slots, self._B = None, None

    # ---------------------------------------------------------------------------------------
    def set_reserve(self, sms: int):
        """SMs (a multiple of 4, 4..28) that the convolution grids leave to the decode = CTA limit of the decode kernels.
        Results do not depend on it; captured graphs are dropped (they hold the old grids)."""
        sms = max(4, min(28, int(sms) // 4 * 4))
        self.model.tuning = (self.model.tuning & ~_abi.TUNE_RESERVE_SMS(7)) | _abi.TUNE_RESERVE_SMS(sms // 4)
        self.params.max_ctas = sms
        self.reserve_sms = sms
        if self._slots is not None:
            torch.cuda.synchronize()
            self._prepared = self.model.prepare(self._B, *self.input_hw)      # the launch configuration carries the tuning word
            for slot in self._slots:
                slot["graphs"] = None

    def calibrate(self, frames_dev, candidates=(8, 12, 16, 24), steps=None, min_ms=150.0):
        """Choose the SM split between the forward and the overlapped decode for THIS workload: the decode's work grows with
        the number of people per frame (peaks x pairs), the forward's does not, so crowded scenes need more than the default
        8 SMs or the decode becomes the longer of the two (measured, 12-16 persons at batch 256: decode 5.3 ms on 8 SMs next to
        a 3.6 ms forward).  Times pipelined steps on ``frames_dev`` [B,1,H,W] (device, representative frames) for every
        candidate -- at least ``min_ms`` of them (or exactly ``steps``) -- and keeps the fastest; returns {sms: ms per step}.
        Every rank of a multi-GPU job must call it (same arguments): the record exchange is part of the step, the ranks
        agree on the slowest rank's time per candidate and therefore on the split.  (4 SMs are not a default candidate: for
        1-6 persons per frame they give the device-resident step +0.5 % but leave the decode 0.78 of the step long, and the
        host-fed pipeline (submit / collect) then runs 1 % SLOWER than with 8 -- profiles/r2_bench_spread.txt.)"""
        B = frames_dev.shape[0]
        slots = self._buffers(B)
        for slot in slots:
            slot["x"].copy_(frames_dev)

        def timed(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                self._run_slot(slots[i % self.NSLOT])
            torch.cuda.current_stream().wait_stream(self.decode_stream)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        def agree(v, op):
            if self.peers is None:
                return v
            import torch.distributed as dist
            t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=op, group=self.peers.group)
            return float(t[0])

        res = {}
        for sms in candidates:
            self.set_reserve(sms)
            t = timed(2 * self.NSLOT)                                  # graph capture + warm-up
            if steps is None:
                import torch.distributed as dist
                n = int(agree(max(8, min_ms / max(t, 1e-3)), dist.ReduceOp.MAX if self.peers is not None else None))
            else:
                n = steps
            import torch.distributed as dist
            res[self.reserve_sms] = agree(timed(n), dist.ReduceOp.MAX if self.peers is not None else None)
        # the fastest split; between candidates within 0.3 % of it the LARGER decode share wins (headroom for frames with more
        # people than the calibration batch: once the decode is the longer half, the step time is the decode's)
        best = min(res.values())
        self.set_reserve(max(k for k, v in res.items() if v <= 1.003 * best))
        return res

    # the two halves of a step, as plain launch sequences on the current stream
    def _launch_forward(self, slot):
        self.model.forward_into(slot["x"], slot["maps"], self._prepared)

    def _launch_decode(self, slot, d2h=True):
        paf, heat, depth = slot["maps"][3:6]
        if self.inject is not None:
            heat, paf, depth = self.inject
        push = None if self.peers is None else self.peers.push_args(slot["index"])
        self.backend.decode_device(heat, paf, depth, self.params, slot["out"], push=push)
        if self.peers is not None:
            self.peers.wait_arrivals(slot["index"])          # one-warp kernel: all ranks' records of this step have landed
        if d2h:
            slot["host"].copy_(slot["out"]["_records"], non_blocking=True)

    def _capture(self, slot):
        """Capture the slot's forward and decode sequences.  Warm-up launches first (lazy module loading, attribute
        calls), then capture on side streams as torch requires."""
        if self.inject is not None:
            raise RuntimeError("inject is a decode-only test hook: construct the estimator with use_graphs=False")
        s = self._capture_stream          # one side stream for all captures (the forward keeps branch streams per caller stream)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._launch_forward(slot)
            self._launch_decode(slot)
        s.synchronize()
        if self.peers is not None:
            self.peers.barrier()          # the warm-up pushed one record set: every rank is past it before the real ones
        gf, gd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(gf, stream=s, capture_error_mode="thread_local"):
            self._launch_forward(slot)
        with torch.cuda.graph(gd, stream=self._capture_stream_dec, capture_error_mode="thread_local"):
            self._launch_decode(slot)
        slot["graphs"] = (gf, gd)

    def _run_slot(self, slot, evs=None):
        """Enqueue forward (current stream) and decode (decode stream) of a slot whose input buffer is (or will be, by
        stream order) filled.  evs: optional 4 timing events [fwd start, fwd end, decode start, decode end]."""
        main, ds = torch.cuda.current_stream(), self.decode_stream
        graphs = self.use_graphs and self.inject is None
        if graphs and slot["graphs"] is None:
            self._capture(slot)
        # the decode of this slot's previous use must have finished reading the maps before the forward rewrites them
        main.wait_event(slot["done"])
        if evs is not None:
            evs[0].record(main)
        if graphs:
            slot["graphs"][0].replay()
        else:
            self._launch_forward(slot)
        if evs is not None:
            evs[1].record(main)
        slot["fwd_done"].record(main)
        ds.wait_event(slot["fwd_done"])
        with torch.cuda.stream(ds):
            if evs is not None:
                evs[2].record(ds)
            if graphs:
                slot["graphs"][1].replay()
            else:
                self._launch_decode(slot)
            if evs is not None:
                evs[3].record(ds)
            slot["done"].record(ds)

    # ---------------------------------------------------------------------------------------
    def infer_device(self, x_dev, evs=None):
        """x_dev [B,1,H,W] fp32 CUDA -> the slot's dict of DEVICE record tensors (no synchronisation; wait on
        ``out["_ready"]`` before reading them from another stream).  The frames are copied into the slot's input buffer
        (device to device) unless ``x_dev`` is a slot buffer returned by ``slot_input``."""
        B = x_dev.shape[0]
        slots = self._buffers(B)
        slot = next((s for s in slots if s["x"].data_ptr() == x_dev.data_ptr()), None)
        if slot is None:
            slot = slots[self._next % self.NSLOT]
            self._next += 1
            torch.cuda.current_stream().wait_event(slot["done"])
            slot["x"].copy_(x_dev, non_blocking=True)
        self._run_slot(slot, evs)
        slot["out"]["_ready"] = slot["done"]
        return slot["out"]

    def slot_input(self, i, B):
        """The device input buffer of slot i (fill it, then pass it to ``infer_device``: no staging copy)."""
        return self._buffers(B)[i % self.NSLOT]["x"]

    def submit(self, frames):
        """Asynchronous half of ``infer``: enqueue H2D (copy stream) -> forward -> decode -> D2H for one batch of HOST
        frames and return a ticket.  Up to NSLOT batches may be in flight; ``collect`` them in submission order."""
        x = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames, np.float32))
        B = x.shape[0]
        slot = self._buffers(B)[self._next % self.NSLOT]
        if slot["busy"]:
            raise RuntimeError("collect() the oldest batch before submitting a %dth one" % (self.NSLOT + 1))
        self._copy_stream.wait_event(slot["done"])          # the previous user of this slot has finished with x / host
        with torch.cuda.stream(self._copy_stream):
            slot["x"].copy_(x, non_blocking=True)
            slot["h2d"].record(self._copy_stream)
        torch.cuda.current_stream().wait_event(slot["h2d"])
        self._run_slot(slot)
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

    def gathered(self, ticket):
        """Multi-GPU: DEVICE views of ALL ranks' records of a collected batch (rank r's frames at [r*B, (r+1)*B))."""
        slot, B = ticket
        return unpack_records(self.peers.gathered(slot["index"]), slot["out"]["_layout"], B, self.peers.world)

    def infer(self, frames):
        """The user-facing call: ``frames`` [B,1,H,W] fp32 on the HOST (NumPy array or, to avoid a staging copy,
        a pinned torch tensor) -> dict of NumPy pose records.  Includes H2D of the frames and D2H of the records."""
        return self.collect(self.submit(frames))

    def h2d_bytes(self, B):
        return B * self.input_hw[0] * self.input_hw[1] * 4

    def d2h_bytes(self, B):
        return int(self._buffers(B)[0]["out"]["_records"].numel())


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


def gather_ap_rows(conf, labels, n_gt, group=None):
    """Multi-GPU AP (SURVEY.md 5 iii): every rank has run the GT assignment (popnet_eval_map_assign) on its contiguous
    shard of frames and holds one (score[K], label[K]) row per PREDICTED human of the shard plus its nGT[K] counts; the AP
    of the whole set needs all rows in one list (the reference appends them frame by frame, util/eval_mAP.py:132-155,
    and sorts once per joint, :160-191).  Row counts differ per rank: ONE all-gather of the counts, ONE all-gather of the
    rows packed as bytes and padded to the longest shard, one all-reduce of nGT.  Returns (conf[R, K] float64,
    labels[R, K] int32, n_gt[K] int64) with rank r's rows before rank r+1's -- the single-process row order, so the AP tail
    (NumPy or popnet_eval_ap) sees exactly the single-process input.  NCCL on CUDA tensors, gloo on CPU tensors."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    conf = np.ascontiguousarray(np.asarray(conf, np.float64))
    labels = np.ascontiguousarray(np.asarray(labels).astype(np.int32))
    if conf.shape != labels.shape or conf.ndim != 2:
        raise ValueError("conf and labels must both be [rows, K]")
    rows, K = conf.shape
    dev = torch.device("cuda") if dist.get_backend(group) == "nccl" else torch.device("cpu")
    counts = torch.zeros((world,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([rows], dtype=torch.int64, device=dev), group=group)
    counts = counts.cpu().numpy()
    per = int(counts.max()) * K * 12                                  # 8 B score + 4 B label per (row, joint)
    mine = np.zeros((per,), np.uint8)
    mine[:rows * K * 8] = conf.view(np.uint8).reshape(-1)
    mine[int(counts.max()) * K * 8:int(counts.max()) * K * 8 + rows * K * 4] = labels.view(np.uint8).reshape(-1)
    full = torch.empty((world * per,), dtype=torch.uint8, device=dev)
    if per:
        dist.all_gather_into_tensor(full, torch.from_numpy(mine).to(dev), group=group)
    full = full.cpu().numpy().reshape(world, per)
    cm = int(counts.max())
    confs = [full[r, :cm * K * 8].view(np.float64).reshape(cm, K)[:int(counts[r])] for r in range(world)]
    labs = [full[r, cm * K * 8:].view(np.int32).reshape(cm, K)[:int(counts[r])] for r in range(world)]
    tot = reduce_counts({"n_gt": np.asarray(n_gt, np.int64)} if dev.type == "cpu"
                        else {"n_gt": torch.as_tensor(np.asarray(n_gt, np.int64), device=dev)}, group)["n_gt"]
    return np.concatenate(confs, 0), np.concatenate(labs, 0), tot.cpu().numpy()


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
