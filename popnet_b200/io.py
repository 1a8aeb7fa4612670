"""Wire formats of the MP-3DHP release ("next" row 2 of SURVEY.md 8(f)): labels.json, *_results.json / eval_data.json.

  parse_gt_labels          main_evaluate_mp_human_3D.py:20-37
  evaluate_mp_human_3d     the body of main_evaluate_mp_human_3D.py:40-99 (the four published metrics)
  save_eval_data           the dump of evaluate/evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:398-409
"""
from __future__ import annotations

import json

import numpy as np

from . import evaluate
from .topology import NUM_JOINTS, get_keypoints


def parse_gt_labels(anno_dic):
    """{image_id: [{'2d_joints': [K][2], '3d_joints': [K][3], ...}], 'intrinsics': ...} -> (gt2d, gt3d) ragged lists."""
    gt2d, gt3d = [], []
    for key, anns in anno_dic.items():
        if key == "intrinsics":
            continue
        gt2d.append([a["2d_joints"] for a in anns])
        gt3d.append([a["3d_joints"] for a in anns])
    return gt2d, gt3d


def load_results(res_file, aligned=None):
    """Returns (pred2d, pred3d, part_conf).  PoP-Net's own result files carry '*_aligned' keys
    (main_evaluate_mp_human_3D.py:45-50); `aligned=None` picks them when present."""
    data = json.load(open(res_file, "r")) if isinstance(res_file, str) else res_file
    use = ("human_pred_set_2d_aligned" in data) if aligned is None else aligned
    suffix = "_aligned" if use else ""
    return data["human_pred_set_2d" + suffix], data["human_pred_set_3d" + suffix], data["human_pred_set_part_conf"]


def evaluate_mp_human_3d(gt, results, *, aligned=None, w_org=480, h_org=512, verbose=False):
    """gt: labels.json path / dict; results: results.json path / dict -> dict of the four metrics
    (PCKh-0.5 2D, PCK 3D @10 cm, mAP 2D, mAP 3D), computed by the CUDA matching kernels."""
    import contextlib
    import io as _io
    anno = json.load(open(gt, "r")) if isinstance(gt, str) else gt
    gt2d, gt3d = parse_gt_labels(anno)
    pred2d, pred3d, conf = load_results(results, aligned)
    names = get_keypoints()
    sink = contextlib.nullcontext() if verbose else contextlib.redirect_stdout(_io.StringIO())
    with sink:
        d2, k2 = evaluate.eval_human_dataset_2d_PCKh(pred2d, gt2d, num_joints=NUM_JOINTS, head_id=0, neck_id=1, iou_th=0.5)
        d3, k3 = evaluate.eval_human_dataset_3d(pred2d, gt2d, pred3d, gt3d, num_joints=NUM_JOINTS, dist_th=0.1, iou_th=0.5)
        ap2 = evaluate.eval_ap_mpii_v2(pred2d, [list(c) for c in conf], gt2d, [], head_id=0, neck_id=1, joint_names=names, thresh=0.5)
        ap3 = evaluate.eval_ap_3D(pred3d, [list(c) for c in conf], gt3d, [], joint_names=names, thresh=0.1)
    return {"joint_names": names, "pckh_2d": k2, "avg_2d_error": d2, "pck_3d": k3, "avg_3d_error": d3,
            "ap_2d": ap2, "ap_3d": ap3, "overall": {"pckh_2d": float(np.average(k2)), "pck_3d": float(np.average(k3)),
                                                     "map_2d": float(ap2[-1]), "map_3d": float(ap3[-1])}}


def save_eval_data(path, *, pred2d, pred3d, visibility, part_conf, gt2d=None, gt3d=None):
    """Same keys as the reference's eval_data.json (human_pred_set_2d, ...)."""
    data = {"human_pred_set_2d": pred2d, "human_pred_set_3d": pred3d, "human_pred_set_visibility": visibility,
            "human_pred_set_part_conf": part_conf}
    if gt2d is not None:
        data["human_gt_set_2d"] = gt2d
    if gt3d is not None:
        data["human_gt_set_3d"] = gt3d
    with open(path, "w") as f:
        json.dump(data, f, indent=4)
    return data
