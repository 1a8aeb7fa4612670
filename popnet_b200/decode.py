"""Decode + lift -- host half, with the reference's call signatures.

  paf_to_pose(heatmaps, pafs, config)             third_party_methods/lib/utils/paf_to_pose.py:354-377
  paf_to_human_list(joint_list, assoc)            third_party_methods/lib/utils/common.py:5-32
  decode_frames(...)                              the per-frame loop body of
                                                  evaluate/evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:187-316,
                                                  batched: one call per batch instead of one per frame

All arithmetic (NMS, bicubic refinement, limb scoring, greedy assembly, depth lift, back-projection)
runs in popnet_b200/csrc/decode_kernels.cu; this file only converts between the reference's Python
containers and the flat device records.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .topology import MP3DHP, Camera, DecodeConfig

_backend = None   # object with .decode(heat, paf, depth, params) -> dict of arrays (see _cuda_backend)


def _get_backend():
    global _backend
    if _backend is None:
        from ._cuda_backend import CudaBackend
        _backend = CudaBackend()
    return _backend


def records_to_reference(out, f: int, K: int):
    """One frame of flat decode records -> the reference's (joint_list [N,5], person_to_joint_assoc [P,K+2])."""
    cnt = out["peak_count"][f]
    base = np.concatenate([[0], np.cumsum(cnt)])
    rows = []
    for t in range(K):
        n = int(cnt[t])
        if n:
            xy = out["peak_xy"][f, t, :n].astype(np.float64)
            sc = out["peak_score"][f, t, :n].astype(np.float64)
            ids = base[t] + np.arange(n, dtype=np.float64)
            rows.append(np.column_stack([xy, sc, ids, np.full(n, float(t))]))
    joint_list = np.concatenate(rows, 0) if rows else np.zeros((0, 5), np.float64)
    n = int(out["n_person"][f])
    if n == 0:
        return joint_list, np.array([])          # np.array([]) like the reference (shape (0,))
    pk = out["person_peak"][f, :n, :K].astype(np.float64)
    assoc = np.empty((n, K + 2), np.float64)
    assoc[:, :K] = np.where(pk >= 0, pk + base[:K][None, :], -1.0)
    assoc[:, K] = out["person_score"][f, :n]
    assoc[:, K + 1] = out["person_njoint"][f, :n]
    return joint_list, assoc


def paf_to_pose(heatmaps, pafs, config):
    """Same contract as the reference: HWC float32 maps of ONE frame -> (joint_list, person_to_joint_assoc)."""
    cfg = DecodeConfig.from_cfg(config)
    heat = np.ascontiguousarray(np.transpose(np.asarray(heatmaps, np.float32), (2, 0, 1)))[None]
    paf = np.ascontiguousarray(np.transpose(np.asarray(pafs, np.float32), (2, 0, 1)))[None]
    # any H x W like the reference: the grid comes from the maps themselves (rows, columns)
    params = _abi.make_decode_params(cfg, MP3DHP, input_size=heat.shape[2] * cfg.downsample, grid_hw=heat.shape[2:4])
    out = _get_backend().decode(heat, paf, None, params)
    _raise_on_overflow(out["flags"])
    return records_to_reference(out, 0, cfg.num_keypoints)


def paf_to_human_list(joint_list, person_to_joint_assoc):
    """common.py:5-32 (pure container reshaping; kept on the host)."""
    humans, visibility, conf_vec = [], [], []
    for human in person_to_joint_assoc:
        idx = human[:-2].astype(int)
        humans.append([[-1, -1] if i < 0 else joint_list[i, :2].tolist() for i in idx])
        conf_vec.append([0 if i < 0 else float(joint_list[i, 2]) for i in idx])
        visibility.append((idx >= 0).astype(int).tolist())
    return humans, visibility, conf_vec


def _radius(radius):
    """The reference takes any number and truncates `center -/+ radius` (common.py:279-282); integral radii 0..5 run on
    the device (windows of up to 121 cells), anything else is rejected loudly."""
    r = int(radius)
    if r != radius or not 0 <= r <= 5:
        raise ValueError("radius must be an integer in 0..5 (got %r)" % (radius,))
    return r


def retrieve_depth_heat_weighted(center, depthmap, heatmap, radius=1):
    """lib/utils/common.py:272-293: heat-weighted depth in the clipped (2*radius+1)^2 window around ``center`` = (x, y) grid cell.
    ``depthmap`` / ``heatmap`` are single [gh, gw] fp32 maps (already de-normalised depth, as at the reference's call
    site, ...mpreal_ablation.py:212-215).  Returns np.float32.  (The batched decode does this on the device for every
    assembled joint; this entry point exists for callers that use the helper on its own.)"""
    z = _get_backend().lift_depth(np.asarray(heatmap, np.float32)[None], np.asarray(depthmap, np.float32)[None],
                                  np.array([[0, int(center[0]), int(center[1])]], np.int32), radius=_radius(radius))
    return np.float32(z[0])


def retrieve_depth_weighted(center, depthmap, radius=1):
    """lib/utils/common.py:251-269: plain mean of the clipped 3x3 depth window (np.mean of an fp32 window)."""
    z = _get_backend().lift_depth(None, np.asarray(depthmap, np.float32)[None],
                                  np.array([[0, int(center[0]), int(center[1])]], np.int32), mode=_abi.LIFT_MEAN,
                                  radius=_radius(radius))
    return np.float32(z[0])


def retrieve_depth_heat_max(center, depthmap, heatmap, radius=1):
    """lib/utils/common.py:296-318: depth at the (first, row-major) maximum of the heat-map inside the clipped 3x3
    window; like the reference, negative heat values count as 0."""
    z = _get_backend().lift_depth(np.asarray(heatmap, np.float32)[None], np.asarray(depthmap, np.float32)[None],
                                  np.array([[0, int(center[0]), int(center[1])]], np.int32), mode=_abi.LIFT_HEAT_MAX,
                                  radius=_radius(radius))
    return np.float32(z[0])


def _raise_on_overflow(flags):
    bad = np.nonzero(np.asarray(flags))[0]
    if len(bad):
        raise OverflowError("decode capacity exceeded in frame(s) %s (flags %s): more than %d peaks per joint "
                            "type or %d persons" % (bad[:8].tolist(), np.asarray(flags)[bad[:8]].tolist(),
                                                    _abi.MAX_PEAKS, _abi.MAX_PERSONS))


def records_to_lists(out, K: int):
    """Flat records of a batch -> the four ragged lists the reference's eval scripts accumulate
    (human_pred_set_2d, human_pred_set_3d, human_pred_set_visibility, human_pred_set_part_conf)."""
    n = np.asarray(out["n_person"])
    p2, p3, vis, conf = [], [], [], []
    pose2d, pose3d, pconf, ppk = (np.asarray(out[k]) for k in ("pose2d", "pose3d", "pose_conf", "person_peak"))
    for f in range(len(n)):
        m = int(n[f])
        p2.append(pose2d[f, :m, :K].tolist())
        p3.append(pose3d[f, :m, :K].tolist())
        conf.append(pconf[f, :m, :K].tolist())
        vis.append((ppk[f, :m, :K] >= 0).astype(int).tolist())
    return p2, p3, vis, conf


def records_to_packed(out, K: int):
    """Flat records of a batch -> (pred2d, pred3d, conf) as ``evaluate.Packed`` CSR arrays: the evaluator's inputs without
    the detour through Python lists (the reference accumulates lists, ...mpreal_ablation.py:266-274, and dumps them to
    JSON; ``records_to_lists`` produces those)."""
    from .evaluate import Packed
    n = np.asarray(out["n_person"]).astype(np.int64)
    off = np.zeros(len(n) + 1, np.int32)
    np.cumsum(n, out=off[1:])
    M = np.asarray(out["pose2d"]).shape[1]
    keep = (np.arange(M)[None, :] < n[:, None])                      # [B, M] valid person slots, frame-major order
    sel = lambda a: np.ascontiguousarray(np.asarray(a)[:, :, :K][keep], np.float64)
    return Packed(sel(out["pose2d"]), off), Packed(sel(out["pose3d"]), off), Packed(sel(out["pose_conf"]), off)


def decode_frames(heat, paf, depth, config=None, camera: Camera = MP3DHP, *, input_size: int = 224,
                  strict: bool = True):
    """Batched decode + lift of channel-major maps (NumPy or CUDA tensors):
    heat [B,K+1,g,g], paf [B,2L,g,g], depth [B,K,g,g] -> dict of flat records (NumPy)."""
    cfg = DecodeConfig.from_cfg(config) if config is not None else DecodeConfig()
    params = _abi.make_decode_params(cfg, camera, input_size=input_size, grid_hw=tuple(heat.shape[2:4]),
                                     depth_channels=0 if depth is None else int(depth.shape[1]))
    out = _get_backend().decode(heat, paf, depth, params)
    if strict:
        _raise_on_overflow(out["flags"])
    return out
