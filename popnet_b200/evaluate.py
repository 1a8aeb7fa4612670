"""Best-match PCK and MPII-style mAP evaluation -- host half.

Same call signatures and return values as the reference evaluator
  util/eval_pck.py:20   eval_human_dataset_2d
  util/eval_pck.py:80   eval_human_dataset_2d_PCKh
  util/eval_pck.py:313  eval_human_dataset_3d
  util/eval_mAP.py:272  eval_ap_mpii_v2
  util/eval_mAP.py:335  eval_ap_3D
(driver: main_evaluate_mp_human_3D.py:58-99; identical copies in
third_party_methods/evaluate/eval_pose_mp.py and eval_ap_mpii.py).

What runs where
  device  (popnet_eval_pck / popnet_eval_map_assign, csrc/eval_kernels.cu): bounding boxes, IoU
          matrix, GT->pred / pred->GT matching, per-joint distances, hit flags, labels, all integer
          counters -- one warp per frame, fp64, rounding matched to NumPy/BLAS.
  host    (this file): ragged lists -> CSR arrays, the per-GT head size with the reference's own Python
          expression (float.__pow__ is not x*x), and the O(n log n) AP tail
          (sort by confidence, precision/recall, VOC envelope) with the same NumPy calls as the reference
          so that AP values are bit-identical once labels are.

There is no CPU fallback: without the CUDA library every function raises (popnet_b200._lib).
"""
from __future__ import annotations

import numpy as np

from . import _abi

__all__ = ["eval_human_dataset_2d", "eval_human_dataset_2d_PCKh", "eval_human_dataset_3d",
           "eval_ap_mpii_v2", "eval_ap_3D", "eval_ap_3D_sharded", "eval_ap_mpii_v2_sharded", "match_counts", "pack_humans",
           "Packed"]

_backend = None   # object with .pck(arrs, dist_th, iou_th, K) and .map_assign(arrs, thresh, K, D)


def _get_backend():
    global _backend
    if _backend is None:
        from ._cuda_backend import CudaBackend   # raises if the CUDA library / a GPU is missing
        _backend = CudaBackend()
    return _backend


# ---------------------------------------------------------------------------------------------
# packing
# ---------------------------------------------------------------------------------------------
def _packer():
    """popnet_b200/_packlists.so (csrc/packlists.c, built by popnet_b200.build): walks the reference's nested lists at
    list-access speed (np.asarray on them was 85 % of a public evaluate call).  Host code, no fallback."""
    try:
        from . import _packlists
    except ImportError as e:
        from ._lib import PopnetError
        raise PopnetError("popnet_b200/_packlists.so is missing: build it with `python -m popnet_b200.build`") from e
    return _packlists


class Packed:
    """CSR-packed humans (flat [S,K,D] float64, off [N+1] int32): what ``pack_humans`` returns.  Every evaluator entry
    point also ACCEPTS a Packed in place of a ragged list (``io.load_results`` produces them), which skips the packing."""
    __slots__ = ("flat", "off")

    def __init__(self, flat, off):
        self.flat, self.off = flat, off

    def __len__(self):
        return len(self.off) - 1

    def __getitem__(self, f):              # frame f's humans, like indexing the ragged list
        if not 0 <= f < len(self):
            raise IndexError(f)
        return self.flat[self.off[f]:self.off[f + 1]]

    def __iter__(self):
        return (self.flat[self.off[f]:self.off[f + 1]] for f in range(len(self)))


def pack_humans(human_set, K: int, D: int):
    """Ragged [N][n_f][K][D] lists (or a Packed) -> (flat [S,K,D] float64, off [N+1] int32)."""
    if isinstance(human_set, Packed):
        if human_set.flat.shape[1:] != (K, D):
            raise ValueError("every human must be a %d x %d list, got %r" % (K, D, human_set.flat.shape[1:]))
        return human_set.flat, human_set.off
    flat, off = _packer().pack_humans(human_set, K, D)
    return np.frombuffer(flat, np.float64).reshape(-1, K, D), np.frombuffer(off, np.int32)


def _pack_rows(per_frame_rows, K: int, dtype):
    if isinstance(per_frame_rows, Packed):
        flat = np.ascontiguousarray(per_frame_rows.flat, np.float64).reshape(-1, K)
        return flat if dtype == np.float64 else np.ascontiguousarray(flat.astype(dtype))
    flat = np.frombuffer(_packer().pack_rows(per_frame_rows, K), np.float64).reshape(-1, K)
    return flat if dtype == np.float64 else np.ascontiguousarray(flat.astype(dtype))


def _frame_counts(human_set):
    if isinstance(human_set, Packed):
        return np.diff(human_set.off)
    return np.fromiter((len(h) for h in human_set), np.int64, len(human_set))


def _ones_rows(human_set, K):
    """[np.ones((len(g), K)).tolist() for g in human_set] (eval_pck.py:107-111, eval_mAP.py:298-307), built without NumPy"""
    return [[[1.0] * K for _ in range(n)] for n in _frame_counts(human_set).tolist()]


def _head_sizes(humans_gt_set, ind1: int, ind2: int):
    """compute_head_size (eval_pck.py:232-246) / compute_head_size_from_two_joints (eval_mAP.py:26-40),
    evaluated with the reference's expression on the caller's own number types."""
    out = []
    for humans in humans_gt_set:                   # (a Packed iterates frame by frame: the same expression on float64 scalars)
        for human in humans:
            out.append(2 * np.sqrt((human[ind1][0] - human[ind2][0]) ** 2 + (human[ind1][1] - human[ind2][1]) ** 2))
    return out


# ---------------------------------------------------------------------------------------------
# PCK
# ---------------------------------------------------------------------------------------------
def _run_pck(pred2d_set, gt2d_set, pred3d_set, gt3d_set, K, dist_th, iou_th, vis_set, gt_thresh):
    assert len(gt2d_set) == len(pred2d_set)
    pred2d, pred_off = pack_humans(pred2d_set, K, 2)
    gt2d, gt_off = pack_humans(gt2d_set, K, 2)
    arrs = {"pred2d": pred2d, "pred_off": pred_off, "gt2d": gt2d, "gt_off": gt_off}
    if pred3d_set is not None:
        arrs["pred3d"], off3 = pack_humans(pred3d_set, K, 3)
        arrs["gt3d"], goff3 = pack_humans(gt3d_set, K, 3)
        if not (np.array_equal(off3, pred_off) and np.array_equal(goff3, gt_off)):
            raise ValueError("2D and 3D human lists differ in shape")
    vis_all = None
    if vis_set is not None:
        # the reference only walks visibility rows of frames that have GT humans (eval_pck.py:50-58)
        vis_all = _pack_rows([v for v, n in zip(vis_set, _frame_counts(gt2d_set)) if n > 0], K, np.float64)
        if vis_all.shape[0] != gt2d.shape[0]:
            raise ValueError("visibility rows do not match GT humans")
        arrs["gt_vis"] = np.ascontiguousarray((vis_all != 0).astype(np.uint8))
    if gt_thresh is not None:
        arrs["gt_thresh"] = np.ascontiguousarray(np.asarray(gt_thresh, np.float64))
    out = _get_backend().pck(arrs, dist_th=float(dist_th), iou_th=float(iou_th), K=K)
    if np.any(out["status"] != 0):
        # eval_pck.py:441-443 returns an empty bbox array for the frame, :462 then indexes it
        raise IndexError("too many indices for array: a ground-truth human has no valid joint "
                         "(frame %d)" % int(np.nonzero(out["status"])[0][0]))
    samples_cnt = int(gt2d.shape[0])
    return out, vis_all, samples_cnt


def _summarise(out, vis_all, samples_cnt, K, use_vis_denominator):
    d = out["dists"]
    joint_avg_dist, joint_KCP = [], []
    for k in range(K):
        col = d[:, k]
        joint_avg_dist.append(np.average(col[np.where(col >= 0)]))
        hit_cnt = np.int64(out["hit_cnt"][k])
        if use_vis_denominator:
            joint_KCP.append(hit_cnt / np.sum(vis_all[:, k]))
        else:
            joint_KCP.append(hit_cnt / samples_cnt)
    return joint_avg_dist, joint_KCP


def eval_human_dataset_2d(humans_pred_set, humans_gt_set, num_joints=15, dist_th=10.0, iou_th=0.5,
                          human_gt_set_visibility=None):
    """util/eval_pck.py:20-77."""
    out, vis_all, n = _run_pck(humans_pred_set, humans_gt_set, None, None, num_joints, dist_th, iou_th,
                               human_gt_set_visibility, None)
    use_vis = vis_all is not None and vis_all.shape[0] != 0
    return _summarise(out, vis_all, n, num_joints, use_vis)


def eval_human_dataset_2d_PCKh(humans_pred_set, humans_gt_set, head_id, neck_id, num_joints=15, h_th=0.5,
                               iou_th=0.5, human_gt_set_visibility=None):
    """util/eval_pck.py:80-154 (per-GT threshold h_th * head size; default visibility = all ones)."""
    assert len(humans_gt_set) == len(humans_pred_set)
    if human_gt_set_visibility is None:
        human_gt_set_visibility = _ones_rows(humans_gt_set, num_joints)
    hsz = _head_sizes(humans_gt_set, head_id, neck_id)
    gt_thresh = [h * h_th for h in hsz]
    out, vis_all, n = _run_pck(humans_pred_set, humans_gt_set, None, None, num_joints, 0.0, iou_th,
                               human_gt_set_visibility, gt_thresh)
    use_vis = vis_all.shape[0] != 0
    return _summarise(out, vis_all, n, num_joints, use_vis)


def _head_sizes_from_rect(head_sz_set, humans_gt_set=None, SC_BIAS=0.6):
    """compute_head_size_from_rect (eval_pck.py:249-263 / eval_mAP.py:43-57) with the reference's expression on the caller's
    number types; frames without GT humans are skipped when `humans_gt_set` is given (eval_pck.py:199-203)."""
    out = []
    for i, rects in enumerate(head_sz_set):
        if humans_gt_set is not None and len(humans_gt_set[i]) == 0:
            continue
        for rect in rects:
            out.append(np.sqrt((rect[2] - rect[0]) ** 2 + (rect[3] - rect[1]) ** 2) * SC_BIAS)
    return out


def eval_human_dataset_2d_PCKh_rect(humans_pred_set, humans_gt_set, head_sz_set, num_joints=15, h_th=0.5, iou_th=0.5,
                                    human_gt_set_visibility=None):
    """util/eval_pck.py:157-229: PCKh with the head size taken from annotated head rectangles (x1, y1, x2, y2),
    hsz = 0.6 * diagonal."""
    assert len(humans_gt_set) == len(humans_pred_set)
    if human_gt_set_visibility is None:
        human_gt_set_visibility = _ones_rows(humans_gt_set, num_joints)
    hsz = _head_sizes_from_rect(head_sz_set, humans_gt_set)
    gt_thresh = [h * h_th for h in hsz]
    out, vis_all, n = _run_pck(humans_pred_set, humans_gt_set, None, None, num_joints, 0.0, iou_th,
                               human_gt_set_visibility, gt_thresh)
    use_vis = vis_all.shape[0] != 0
    return _summarise(out, vis_all, n, num_joints, use_vis)


def eval_human_dataset_3d(humans_pred_set_2d, humans_gt_set_2d, humans_pred_set_3d, humans_gt_set_3d,
                          num_joints=15, dist_th=0.1, iou_th=0.5, human_gt_set_visibility=None):
    """util/eval_pck.py:313-374."""
    out, vis_all, n = _run_pck(humans_pred_set_2d, humans_gt_set_2d, humans_pred_set_3d, humans_gt_set_3d,
                               num_joints, dist_th, iou_th, human_gt_set_visibility, None)
    return _summarise(out, vis_all, n, num_joints, human_gt_set_visibility is not None)


def match_counts(humans_pred_set, humans_gt_set, *, pred3d=None, gt3d=None, num_joints=15, dist_th=0.1,
                 iou_th=0.5, gt_thresh=None, human_gt_set_visibility=None):
    """The integer contract on its own: dict(hit_cnt[K], valid_cnt[K], samples_cnt, dists, matched_pred)."""
    out, _, n = _run_pck(humans_pred_set, humans_gt_set, pred3d, gt3d, num_joints, dist_th, iou_th,
                         human_gt_set_visibility, gt_thresh)
    return {"hit_cnt": out["hit_cnt"].astype(np.int64), "valid_cnt": out["valid_cnt"].astype(np.int64),
            "samples_cnt": n, "dists": out["dists"], "matched_pred": out["matched_pred"], "hit": out["hit"]}


# ---------------------------------------------------------------------------------------------
# mAP
# ---------------------------------------------------------------------------------------------
def _get_rpc(class_margin, true_labels, totalpos):
    """getRPC (eval_mAP.py:160-191) with the loop replaced by a cumulative sum (same quotients)."""
    class_margin = np.array(class_margin)
    true_labels = np.array(true_labels)
    sortidx = np.flip(np.argsort(class_margin))
    sorted_labels = true_labels[sortidx]
    npos = np.cumsum(sorted_labels == 1)
    precision = npos / np.arange(1, len(sorted_labels) + 1)
    recall = npos / totalpos
    return precision.astype(np.float64), recall.astype(np.float64)


def _voc_ap(recall, precision):
    """VOCap (eval_mAP.py:194-207)."""
    vecN = len(recall) + 2
    mrec = np.zeros(vecN)
    mrec[1:-1] = recall
    mrec[-1] = 1.
    mpre = np.zeros(vecN)
    mpre[1:-1] = precision
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    indices = np.where((mrec[1:] - mrec[:-1]) > 0)[0] + 1
    return np.sum((mrec[indices] - mrec[indices - 1]) * mpre[indices])


def _assign(humans_pred_set, conf_pred_set, humans_gt_set, gt_visibility_set, ref_dist_set, K, D, thresh):
    pred, pred_off = pack_humans(humans_pred_set, K, D)
    gt, gt_off = pack_humans(humans_gt_set, K, D)
    if np.any((np.diff(pred_off) > 0) & (np.diff(gt_off) == 0)):
        # eval_mAP.py:124 np.argmax over an empty GT axis
        raise ValueError("attempt to get argmax of an empty sequence")
    conf = np.ones((pred.shape[0], K)) if conf_pred_set is _ALL_ONES else _pack_rows(conf_pred_set, K, np.float64)
    if gt_visibility_set is _ALL_ONES:
        vis8 = np.ones((gt.shape[0], K), np.uint8)
    else:
        vis8 = np.ascontiguousarray((_pack_rows(gt_visibility_set, K, np.float64) > 0).astype(np.uint8))
    if ref_dist_set is _ALL_ONES:
        ref = np.ones(gt.shape[0], np.float64)
    else:
        ref = np.ascontiguousarray(np.asarray([r for fr in ref_dist_set for r in fr], np.float64))
    arrs = {"pred": pred, "pred_off": pred_off, "gt": gt, "gt_off": gt_off, "ref_dist": ref, "gt_vis": vis8}
    out = _get_backend().map_assign(arrs, thresh=float(thresh), K=K, D=D)
    return out, conf


#: "device" (default) = popnet_eval_ap: per-joint sort by (score descending, prediction index ascending) -- one of the
#: orders the reference's unstable np.argsort can produce --, true-positive scan, VOC envelope; float64 sums in a
#: different order than NumPy's pairwise sum: AP agrees with the reference to ~1e-12 (tests assert 1e-9).
#: "numpy" = the reference's own NumPy calls for the sort / scan / envelope (util/eval_mAP.py:160-207) on the device's labels.
AP_TAIL = "device"
_ALL_ONES = object()          # sentinel: "this input is the reference's all-ones default" -- nothing to pack


def _ap_from_labels(out, conf, joint_names):
    K = len(joint_names)
    if AP_TAIL == "device" and hasattr(_get_backend(), "ap_tail"):       # (the CPU test seam has no device tail)
        ap = np.asarray(_get_backend().ap_tail(conf, out["labels"], out["n_gt"]), np.float64)
    else:
        ap = np.zeros(K + 1)
        for j in range(K):
            scores = conf[:, j]
            labels = out["labels"][:, j].astype(np.int64)
            precision, recall = _get_rpc(scores, labels, np.float64(out["n_gt"][j]))
            ap[j] = _voc_ap(recall, precision) * 100
        ap[-1] = np.mean(ap[:-1])
    for j, name in enumerate(joint_names):
        print('    {},  AP: {:03f}'.format(name, ap[j]))
    print('\n     Overall: AP: {:03f}\n'.format(ap[-1]))
    return ap


def _fill_defaults(humans_pred_set, conf_pred_set, humans_gt_set, gt_visibility_set, K):
    """eval_mAP.py:297-307: empty lists are filled in place with all-ones rows (the caller's lists are mutated, as in the
    reference).  Returns the (conf, visibility) inputs for the packer: the sentinel when we just wrote the ones ourselves."""
    vis, conf = gt_visibility_set, conf_pred_set
    if len(gt_visibility_set) == 0:
        gt_visibility_set.extend(_ones_rows(humans_gt_set, K))
        vis = _ALL_ONES
    if len(conf_pred_set) == 0:
        conf_pred_set.extend(_ones_rows(humans_pred_set, K))
        conf = _ALL_ONES
    return conf, vis


def eval_ap_mpii_v2(humans_pred_set, conf_pred_set, humans_gt_set, gt_visibility_set, head_id, neck_id,
                    joint_names, thresh=0.5, _return_counts=False):
    """util/eval_mAP.py:272-332: 2D AP under the PCKh rule, reference distance 2*|head - neck| of the GT."""
    print('2D evaluation in AP evaluation under PCKh-{:01f} rule ...'.format(thresh))
    assert len(humans_gt_set) == len(humans_pred_set)
    K = len(joint_names)
    ref_dist_set = [_head_sizes([g], head_id, neck_id) for g in humans_gt_set]
    conf_in, vis_in = _fill_defaults(humans_pred_set, conf_pred_set, humans_gt_set, gt_visibility_set, K)
    out, conf = _assign(humans_pred_set, conf_in, humans_gt_set, vis_in, ref_dist_set, K, 2, thresh)
    ap = _ap_from_labels(out, conf, joint_names)
    return (ap, out) if _return_counts else ap


def eval_ap_mpii(humans_pred_set, conf_pred_set, humans_gt_set, gt_visibility_set, head_sz_set, joint_names, thresh=0.5,
                 _return_counts=False):
    """util/eval_mAP.py:210-269: 2D AP under the PCKh rule with the head size from annotated head rectangles."""
    print('2D evaluation in AP evaluation under PCKh-{:01f} rule ...'.format(thresh))
    assert len(humans_gt_set) == len(humans_pred_set)
    K = len(joint_names)
    ref_dist_set = [_head_sizes_from_rect([head_sz_set[i]]) for i in range(len(humans_gt_set))]
    conf_in, vis_in = _fill_defaults(humans_pred_set, conf_pred_set, humans_gt_set, gt_visibility_set, K)
    out, conf = _assign(humans_pred_set, conf_in, humans_gt_set, vis_in, ref_dist_set, K, 2, thresh)
    ap = _ap_from_labels(out, conf, joint_names)
    return (ap, out) if _return_counts else ap


def eval_ap_3D(humans_pred_set, conf_pred_set, humans_gt_set, gt_visibility_set, joint_names, thresh=0.1,
               _return_counts=False):
    """util/eval_mAP.py:335-395: 3D AP under the 10 cm rule (reference distance 1)."""
    print('3D evaluation in AP under {:01f} meter rule ...'.format(thresh))
    assert len(humans_gt_set) == len(humans_pred_set)
    K = len(joint_names)
    conf_in, vis_in = _fill_defaults(humans_pred_set, conf_pred_set, humans_gt_set, gt_visibility_set, K)
    out, conf = _assign(humans_pred_set, conf_in, humans_gt_set, vis_in, _ALL_ONES, K, 3, thresh)
    ap = _ap_from_labels(out, conf, joint_names)
    return (ap, out) if _return_counts else ap


def eval_ap_3D_sharded(humans_pred_shard, conf_pred_shard, humans_gt_shard, gt_visibility_shard, joint_names, thresh=0.1,
                       group=None):
    """eval_ap_3D over a data set sharded across the ranks of a process group (contiguous frame shards,
    pipeline.shard): GT assignment on each rank's shard (device), then ``pipeline.gather_ap_rows`` and the AP tail on the
    gathered rows.  Every rank returns the AP vector of the WHOLE set, equal to the single-process eval_ap_3D's
    (util/eval_mAP.py:335-395 on the concatenated lists)."""
    from . import pipeline
    K = len(joint_names)
    conf_in, vis_in = _fill_defaults(humans_pred_shard, conf_pred_shard, humans_gt_shard, gt_visibility_shard, K)
    out, conf = _assign(humans_pred_shard, conf_in, humans_gt_shard, vis_in, _ALL_ONES, K, 3, thresh)
    conf_all, labels_all, n_gt = pipeline.gather_ap_rows(conf, out["labels"], out["n_gt"], group)
    return _ap_from_labels({"labels": labels_all, "n_gt": n_gt}, conf_all, joint_names)


def eval_ap_mpii_v2_sharded(humans_pred_shard, conf_pred_shard, humans_gt_shard, gt_visibility_shard, head_id, neck_id,
                            joint_names, thresh=0.5, group=None):
    """eval_ap_mpii_v2 (util/eval_mAP.py:272-332) over contiguous frame shards; see eval_ap_3D_sharded."""
    from . import pipeline
    K = len(joint_names)
    ref_dist_set = [_head_sizes([g], head_id, neck_id) for g in humans_gt_shard]
    conf_in, vis_in = _fill_defaults(humans_pred_shard, conf_pred_shard, humans_gt_shard, gt_visibility_shard, K)
    out, conf = _assign(humans_pred_shard, conf_in, humans_gt_shard, vis_in, ref_dist_set, K, 2, thresh)
    conf_all, labels_all, n_gt = pipeline.gather_ap_rows(conf, out["labels"], out["n_gt"], group)
    return _ap_from_labels({"labels": labels_all, "n_gt": n_gt}, conf_all, joint_names)
