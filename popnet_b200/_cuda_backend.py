"""The one backend of the product: device buffers through torch, compute through libpopnet_b200.so.

torch is plumbing here (allocation, streams, H2D/D2H); every kernel is in popnet_b200/csrc.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi, _lib


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.PopnetError("popnet_b200 needs a CUDA device (no CPU fallback exists)")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_dev(a, dtype=None):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        t = a
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.cuda().contiguous()
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.cuda()


_FIELDS = (  # name, dtype, shape suffix builder -- 8-byte fields first so every view stays aligned
    ("conn_score", torch.float64, lambda K, L, P, M: (L, P)),
    ("person_score", torch.float64, lambda K, L, P, M: (M,)),
    ("pose2d", torch.float64, lambda K, L, P, M: (M, K, 2)),
    ("pose3d", torch.float64, lambda K, L, P, M: (M, K, 3)),
    ("pose_conf", torch.float64, lambda K, L, P, M: (M, K)),
    ("peak_count", torch.int32, lambda K, L, P, M: (K,)),
    ("peak_score", torch.float32, lambda K, L, P, M: (K, P)),
    ("conn_count", torch.int32, lambda K, L, P, M: (L,)),
    ("n_person", torch.int32, lambda K, L, P, M: ()),
    ("person_njoint", torch.int32, lambda K, L, P, M: (M,)),
    ("flags", torch.int32, lambda K, L, P, M: ()),
    ("peak_xy", torch.int16, lambda K, L, P, M: (K, P, 2)),
    ("conn_ij", torch.int16, lambda K, L, P, M: (L, P, 2)),
    ("person_peak", torch.int16, lambda K, L, P, M: (M, K)),
)
#: the fields that leave the device / cross the NVLink fabric, laid out contiguously at the END of the buffer
RECORD_FIELDS = ("person_score", "pose2d", "pose3d", "pose_conf", "n_person", "person_njoint", "flags", "person_peak")


def records_layout(B, params):
    """(byte size, [(name, dtype, shape suffix, offset, nbytes)]) of the pose-record block of a batch."""
    K, L, P, M = params.num_joints, params.num_limbs, params.max_peaks, params.max_persons
    off, lay = 0, []
    for name, dt, shp in _FIELDS:
        if name in RECORD_FIELDS:
            n = B * int(np.prod(shp(K, L, P, M), dtype=np.int64)) * torch.empty((), dtype=dt).element_size()
            lay.append((name, dt, shp(K, L, P, M), off, n))
            off = (off + n + 7) // 8 * 8
    return off, lay


def alloc_decode_out(B, params, device="cuda", records=None):
    """All decode outputs of a batch live in ONE device buffer (field-major: [field][B][...]); the pose-record fields
    sit contiguously at its end so that the multi-GPU exchange (and the D2H copy) is a single transfer of
    ``out["_records"]`` without any packing kernel.  ``records``: optional preallocated uint8 tensor that holds the
    record block instead (the peer-visible buffer of the multi-GPU path)."""
    if records is not None:
        nbytes, lay = records_layout(B, params)
        assert records.numel() >= nbytes and records.dtype == torch.uint8
        out = alloc_decode_out(B, params, device)
        rec = records[:nbytes]
        rec.zero_()
        out["_records"], out["_layout"] = rec, lay
        for name, dt, shp_, o, nb in lay:
            out[name] = rec[o:o + nb].view(dt).reshape((B,) + tuple(shp_))
        return out
    K, L, P, M = params.num_joints, params.num_limbs, params.max_peaks, params.max_persons
    order = [f for f in _FIELDS if f[0] not in RECORD_FIELDS] + [f for f in _FIELDS if f[0] in RECORD_FIELDS]
    sizes, off = [], 0
    rec_start = None
    for name, dt, shp in order:
        if name in RECORD_FIELDS and rec_start is None:
            off = (off + 15) // 16 * 16
            rec_start = off
        n = B * int(np.prod(shp(K, L, P, M), dtype=np.int64)) * torch.empty((), dtype=dt).element_size()
        sizes.append((name, dt, shp(K, L, P, M), off, n))
        off = (off + n + 7) // 8 * 8
    buf = torch.zeros(off, dtype=torch.uint8, device=device)
    out = {"_buffer": buf, "_records": buf[rec_start:], "_layout": [(n_, dt, shp_, o - rec_start, nb) for n_, dt, shp_, o, nb in sizes
                                                                    if n_ in RECORD_FIELDS]}
    for name, dt, shp_, o, nb in sizes:
        out[name] = buf[o:o + nb].view(dt).reshape((B,) + tuple(shp_))
    return out


def check_decode_shapes(heat, paf, depth, params):
    """The C ABI takes raw pointers: the map shapes must agree with the DecodeParams block or the kernels would index
    with the wrong pitch / plane stride.  Raises (like POPNET_ERR_INVALID_ARG) instead of decoding garbage."""
    B, K, L = heat.shape[0], params.num_joints, params.num_limbs
    g = (params.grid_h, params.grid_w)
    dc = params.depth_channels if params.depth_channels > 0 else K
    want = {"heat": (B, K + 1) + g, "paf": (B, 2 * L) + g}
    got = {"heat": tuple(heat.shape), "paf": tuple(paf.shape)}
    if depth is not None:
        want["depth"], got["depth"] = (B, dc) + g, tuple(depth.shape)
    for k in want:
        if want[k] != got[k]:
            raise _lib.PopnetError("popnet_decode: %s has shape %s but the decode parameters (K=%d, L=%d, grid %dx%d, "
                                   "depth_channels=%d) need %s" % (k, got[k], K, L, g[0], g[1], dc, want[k]))
    for k, t in (("heat", heat), ("paf", paf), ("depth", depth)):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda):
            raise _lib.PopnetError("popnet_decode: %s must be a contiguous fp32 CUDA tensor" % k)


class CudaBackend:
    name = "cuda-sm100a"

    def __init__(self):
        _require_cuda()
        self.lib = _lib.get()

    # ------------------------------------------------------------------ evaluator
    def pck_device(self, d, *, dist_th, iou_th, K, out=None):
        """popnet_eval_pck on device-resident CSR tensors; returns device tensors (no synchronisation).  `out` may be
        passed back in to reuse the result buffers."""
        SG = d["gt2d"].shape[0]
        N = d["gt_off"].shape[0] - 1
        if out is None:
            out = {"dists": torch.empty((SG, K), dtype=torch.float64, device="cuda"),
                   "hit": torch.empty((SG, K), dtype=torch.uint8, device="cuda"),
                   "matched_pred": torch.empty((SG,), dtype=torch.int32, device="cuda"),
                   "hit_cnt": torch.empty((K,), dtype=torch.int64, device="cuda"),
                   "valid_cnt": torch.empty((K,), dtype=torch.int64, device="cuda"),
                   "status": torch.zeros((max(N, 1),), dtype=torch.int32, device="cuda")}
        a = _abi.PckArgs(pred2d=_ptr(d["pred2d"]), pred3d=_ptr(d.get("pred3d")), pred_off=_ptr(d["pred_off"]),
                         gt2d=_ptr(d["gt2d"]), gt3d=_ptr(d.get("gt3d")), gt_off=_ptr(d["gt_off"]),
                         gt_vis=_ptr(d.get("gt_vis")), gt_thresh=_ptr(d.get("gt_thresh")),
                         dist_th=dist_th, iou_th=iou_th, num_frames=N, num_joints=K,
                         **{k: _ptr(v) for k, v in out.items()})
        _lib.check(self.lib.popnet_eval_pck(C.byref(a), _stream()), "popnet_eval_pck")
        return out

    def pck(self, arrs, *, dist_th, iou_th, K):
        d = {k: _to_dev(v) for k, v in arrs.items()}
        N = d["gt_off"].shape[0] - 1
        out = self.pck_device(d, dist_th=dist_th, iou_th=iou_th, K=K)
        res = {k: v.cpu().numpy() for k, v in out.items()}
        res["status"] = res["status"][:N]
        return res

    def map_assign_device(self, d, *, thresh, K, D, out=None):
        """popnet_eval_map_assign on device-resident CSR tensors; returns device tensors (no synchronisation)."""
        SP = d["pred"].shape[0]
        N = d["gt_off"].shape[0] - 1
        if out is None:
            out = {"labels": torch.zeros((SP, K), dtype=torch.uint8, device="cuda"),
                   "matched_gt": torch.full((SP,), -1, dtype=torch.int32, device="cuda"),
                   "n_gt": torch.empty((K,), dtype=torch.int64, device="cuda"),
                   "n_pos": torch.empty((K,), dtype=torch.int64, device="cuda")}
        a = _abi.MapArgs(pred=_ptr(d["pred"]), pred_off=_ptr(d["pred_off"]), gt=_ptr(d["gt"]),
                         gt_off=_ptr(d["gt_off"]), gt_vis=_ptr(d.get("gt_vis")), ref_dist=_ptr(d["ref_dist"]),
                         thresh=thresh, num_frames=N, num_joints=K, dim=D,
                         **{k: _ptr(v) for k, v in out.items()})
        _lib.check(self.lib.popnet_eval_map_assign(C.byref(a), _stream()), "popnet_eval_map_assign")
        return out

    def map_assign(self, arrs, *, thresh, K, D):
        d = {k: _to_dev(v) for k, v in arrs.items()}
        return {k: v.cpu().numpy() for k, v in self.map_assign_device(d, thresh=thresh, K=K, D=D).items()}

    def ap_tail_device(self, conf, labels, n_gt):
        """popnet_eval_ap on device tensors: conf [SP,K] f64, labels [SP,K] u8, n_gt [K] i64 -> ap [K+1] f64 (device)."""
        SP, K = conf.shape
        nbytes = self.lib.popnet_eval_ap_workspace_bytes(SP, K)
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device="cuda")
        ap = torch.empty((K + 1,), dtype=torch.float64, device="cuda")
        a = _abi.ApArgs(conf=_ptr(conf), labels=_ptr(labels), n_gt=_ptr(n_gt), num_preds=SP, num_joints=K, ap=_ptr(ap),
                        workspace=_ptr(ws), workspace_bytes=nbytes)
        _lib.check(self.lib.popnet_eval_ap(C.byref(a), _stream()), "popnet_eval_ap")
        ws.record_stream(torch.cuda.current_stream())
        return ap

    def ap_tail(self, conf, labels, n_gt):
        """Host arrays in, ap [K+1] (NumPy) out."""
        c = _to_dev(np.ascontiguousarray(conf, np.float64))
        l = _to_dev(np.ascontiguousarray(labels, np.uint8))
        g = _to_dev(np.ascontiguousarray(n_gt, np.int64))
        return self.ap_tail_device(c, l, g).cpu().numpy()

    # ------------------------------------------------------------------ decode
    def decode_device(self, heat, paf, depth, params, out=None, push=None):
        """Device tensors in, device tensors out (no synchronisation); `out` buffers may be reused.  push: optional
        _abi.PeerPush block (multi-GPU: the assembly kernel also stores every record value into the peers' gather buffers)."""
        B = heat.shape[0]
        check_decode_shapes(heat, paf, depth, params)
        if out is None:
            out = alloc_decode_out(B, params)
        o = _abi.DecodeOut(**{k: _ptr(v) for k, v in out.items() if not k.startswith("_")})
        if push is None:
            _lib.check(self.lib.popnet_decode(_ptr(heat), _ptr(paf), _ptr(depth), B, C.byref(params), C.byref(o),
                                              _stream()), "popnet_decode")
        else:
            _lib.check(self.lib.popnet_decode_push(_ptr(heat), _ptr(paf), _ptr(depth), B, C.byref(params), C.byref(o),
                                                   C.byref(push), _stream()), "popnet_decode_push")
        return out

    def decode(self, heat, paf, depth, params):
        """NumPy (or torch) maps in, dict of NumPy arrays out -- same keys/strides as the C oracle."""
        h, p_, d = _to_dev(heat, torch.float32), _to_dev(paf, torch.float32), _to_dev(depth, torch.float32)
        out = self.decode_device(h, p_, d, params)
        res = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
        res["flags"] = res["flags"].view(np.uint32)
        return res

    def lift_depth(self, heat, depth, queries, depth_mean=0.0, depth_std=1.0, mode=_abi.LIFT_HEAT_WEIGHTED, radius=1):
        """heat/depth: [planes, gh, gw] fp32 (heat may be None for LIFT_MEAN); queries [n,3] int32 (plane, cx, cy)
        -> fp32 [n] (NumPy).  radius: half-width of the clipped window (0..5)."""
        d = _to_dev(depth, torch.float32)
        h = _to_dev(heat, torch.float32) if heat is not None else None
        q = _to_dev(np.ascontiguousarray(queries, np.int32))
        n = q.shape[0]
        out = torch.empty((n,), dtype=torch.float32, device="cuda")
        _lib.check(self.lib.popnet_lift_depth_window(_ptr(h), _ptr(d), _ptr(q), n, d.shape[-2], d.shape[-1], float(depth_mean),
                                                     float(depth_std), int(mode), int(radius), _ptr(out), _stream()),
                   "popnet_lift_depth_window")
        return out.cpu().numpy()

    def preprocess_depth(self, frames, dst_hw, depth_max, depth_mean, depth_std):
        """frames [B, H, W] fp32 metres (NumPy or CUDA) -> CUDA tensor [B, 1, dst_h, dst_w], normalised."""
        src = _to_dev(frames, torch.float32)
        B, H, W = src.shape
        dst = torch.empty((B, 1, dst_hw[0], dst_hw[1]), dtype=torch.float32, device="cuda")
        _lib.check(self.lib.popnet_preprocess_depth(_ptr(src), B, H, W, _ptr(dst), dst_hw[0], dst_hw[1], float(depth_max),
                                                    float(depth_mean), float(depth_std), _stream()), "popnet_preprocess_depth")
        return dst
