"""ORACLE (test infrastructure): ctypes loader for oracle/_ref/libpopnet_oracle.so (oracle/Makefile).

Host-pointer twins of the CUDA entry points, on NumPy arrays.  Only tests/, smoke() and bench.py's
CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from popnet_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libpopnet_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "popnet_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "popnet_b200.h")          # the oracle shares the ABI's structs
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        l = C.CDLL(_SO)
        l.oracle_decode.restype = C.c_int
        l.oracle_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.POINTER(_abi.DecodeParams), C.POINTER(_abi.DecodeOut)]
        l.oracle_eval_pck.restype = C.c_int
        l.oracle_eval_pck.argtypes = [C.POINTER(_abi.PckArgs)]
        l.oracle_eval_map_assign.restype = C.c_int
        l.oracle_eval_map_assign.argtypes = [C.POINTER(_abi.MapArgs)]
        l.oracle_bicubic_upsample.restype = None
        l.oracle_bicubic_upsample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        l.oracle_fma_vec.restype = None
        l.oracle_fma_vec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.oracle_phase_table.restype = None
        l.oracle_phase_table.argtypes = [C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fma(a, b, c):
    a = np.ascontiguousarray(np.broadcast_to(a, np.broadcast(a, b, c).shape), np.float64)
    b = np.ascontiguousarray(np.broadcast_to(b, a.shape), np.float64)
    c = np.ascontiguousarray(np.broadcast_to(c, a.shape), np.float64)
    out = np.empty_like(a)
    lib().oracle_fma_vec(_p(a), _p(b), _p(c), _p(out), a.size)
    return out


def bicubic_upsample(src):
    src = np.ascontiguousarray(src, np.float32)
    out = np.empty((src.shape[0] * 8, src.shape[1] * 8), np.float32)
    lib().oracle_bicubic_upsample(_p(src), src.shape[0], src.shape[1], _p(out))
    return out


def alloc_decode_out(B, params, xp=np):
    """Host output buffers with the strides popnet_decode documents; returns dict of arrays."""
    K, L, P, M = params.num_joints, params.num_limbs, params.max_peaks, params.max_persons
    return {
        "peak_count": np.zeros((B, K), np.int32),
        "peak_xy": np.full((B, K, P, 2), -1, np.int16),
        "peak_score": np.zeros((B, K, P), np.float32),
        "conn_count": np.zeros((B, L), np.int32),
        "conn_ij": np.full((B, L, P, 2), -1, np.int16),
        "conn_score": np.zeros((B, L, P), np.float64),
        "n_person": np.zeros((B,), np.int32),
        "person_peak": np.full((B, M, K), -1, np.int16),
        "person_score": np.zeros((B, M), np.float64),
        "person_njoint": np.zeros((B, M), np.int32),
        "pose2d": np.full((B, M, K, 2), -1.0, np.float64),
        "pose3d": np.zeros((B, M, K, 3), np.float64),
        "pose_conf": np.zeros((B, M, K), np.float64),
        "flags": np.zeros((B,), np.uint32),
    }


def decode(heat, paf, depth, params):
    """heat [B,K+1,g,g], paf [B,2L,g,g], depth [B,K,g,g] (or None) fp32 -> dict of output arrays."""
    heat = np.ascontiguousarray(heat, np.float32)
    paf = np.ascontiguousarray(paf, np.float32)
    depth = None if depth is None else np.ascontiguousarray(depth, np.float32)
    B = heat.shape[0]
    g = (params.grid_h, params.grid_w)
    dc = params.depth_channels if params.depth_channels > 0 else params.num_joints
    assert heat.shape == (B, params.num_joints + 1) + g and paf.shape == (B, 2 * params.num_limbs) + g, (heat.shape, paf.shape, g)
    assert depth is None or depth.shape == (B, dc) + g, (depth.shape, dc, g)
    bufs = alloc_decode_out(B, params)
    out = _abi.DecodeOut(**{k: _p(v) for k, v in bufs.items()})
    rc = lib().oracle_decode(_p(heat), _p(paf), _p(depth), B, C.byref(params), C.byref(out))
    if rc != 0:
        raise RuntimeError("oracle_decode failed: %s" % _abi.STATUS_NAMES.get(rc, rc))
    return bufs


def eval_pck(arrs, *, dist_th, iou_th, K):
    """arrs: dict with pred2d, pred3d|None, pred_off, gt2d, gt3d|None, gt_off, gt_vis|None, gt_thresh|None."""
    SG = arrs["gt2d"].shape[0]
    N = len(arrs["gt_off"]) - 1
    out = {"dists": np.zeros((SG, K), np.float64), "hit": np.zeros((SG, K), np.uint8),
           "matched_pred": np.zeros((SG,), np.int32), "hit_cnt": np.zeros((K,), np.int64),
           "valid_cnt": np.zeros((K,), np.int64), "status": np.zeros((N,), np.int32)}
    a = _abi.PckArgs(pred2d=_p(arrs["pred2d"]), pred3d=_p(arrs.get("pred3d")), pred_off=_p(arrs["pred_off"]),
                     gt2d=_p(arrs["gt2d"]), gt3d=_p(arrs.get("gt3d")), gt_off=_p(arrs["gt_off"]),
                     gt_vis=_p(arrs.get("gt_vis")), gt_thresh=_p(arrs.get("gt_thresh")),
                     dist_th=dist_th, iou_th=iou_th, num_frames=N, num_joints=K,
                     **{k: _p(v) for k, v in out.items()})
    rc = lib().oracle_eval_pck(C.byref(a))
    assert rc == 0
    return out


def eval_map_assign(arrs, *, thresh, K, D):
    SP = arrs["pred"].shape[0]
    N = len(arrs["gt_off"]) - 1
    out = {"labels": np.zeros((SP, K), np.uint8), "matched_gt": np.zeros((SP,), np.int32),
           "n_gt": np.zeros((K,), np.int64), "n_pos": np.zeros((K,), np.int64)}
    a = _abi.MapArgs(pred=_p(arrs["pred"]), pred_off=_p(arrs["pred_off"]), gt=_p(arrs["gt"]),
                     gt_off=_p(arrs["gt_off"]), gt_vis=_p(arrs.get("gt_vis")), ref_dist=_p(arrs["ref_dist"]),
                     thresh=thresh, num_frames=N, num_joints=K, dim=D,
                     **{k: _p(v) for k, v in out.items()})
    rc = lib().oracle_eval_map_assign(C.byref(a))
    assert rc == 0
    return out
