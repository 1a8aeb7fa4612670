"""Import the UNMODIFIED reference hot path from /root/reference (read-only) with import-time shims.

Harness glue only: stubs for packages the reference imports but never uses on
this path (thop, matplotlib, pylab), the two NumPy aliases removed in NumPy >= 1.24, and a duck-typed
config in place of yacs.  Nothing from the reference is copied; this module is used only by
tests/golden/make_golden.py (run in the build container, where /root/reference exists) and by the
optional ``-m refcheck`` tests.  It is never imported on the GPU box.
"""
import os
import sys
import types
from types import SimpleNamespace as NS

import numpy as np

REFERENCE_ROOT = os.environ.get("POPNET_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "third_party_methods", "lib"))


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Returns a namespace with the reference callables on the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if not hasattr(np, "int"):
        np.int = int          # removed aliases used at common.py:16,27,29 and eval_mAP.py:121
    if not hasattr(np, "float"):
        np.float = float
    _stub("thop", profile=lambda *a, **k: (0, 0), clever_format=lambda *a, **k: None)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot", rcParams={})
    mpl.cm = _stub("matplotlib.cm")
    _stub("pylab", rcParams={})
    for p in (os.path.join(REFERENCE_ROOT, "third_party_methods"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from lib.network.rtpose_light3d import rtpose_light3d
        from lib.utils.paf_to_pose import paf_to_pose, NMS, find_connected_joints, group_limbs_of_same_person
        from lib.utils.common import paf_to_human_list, retrieve_depth_heat_weighted
        from util import eval_pck, eval_mAP
    cfg = NS(MODEL=NS(NUM_KEYPOINTS=15, NUM_LIMBS=14, DOWNSAMPLE=8, IMAGE_SIZE=[224, 224]),
             TEST=NS(THRESH_HEATMAP=0.1, THRESH_PAF=0.05, NUM_INTERMED_PTS_BETWEEN_KEYPOINTS=10))
    _loaded = NS(rtpose_light3d=rtpose_light3d, paf_to_pose=paf_to_pose, NMS=NMS,
                 find_connected_joints=find_connected_joints,
                 group_limbs_of_same_person=group_limbs_of_same_person,
                 paf_to_human_list=paf_to_human_list,
                 retrieve_depth_heat_weighted=retrieve_depth_heat_weighted,
                 eval_pck=eval_pck, eval_mAP=eval_mAP, cfg=cfg)
    return _loaded


def reference_decode_frame(ref, heat_chw, paf_chw, depth_chw, cam):
    """The per-frame body of evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:176-263, calling the
    reference functions exactly as that script does (HWC float32 maps, in-place de-normalisation)."""
    heat = np.ascontiguousarray(np.transpose(heat_chw, (1, 2, 0))).astype(np.float32)
    paf = np.ascontiguousarray(np.transpose(paf_chw, (1, 2, 0))).astype(np.float32)
    posedepth = np.ascontiguousarray(np.transpose(depth_chw, (1, 2, 0))).astype(np.float32)
    posedepth *= cam.depth_std
    posedepth += cam.depth_mean
    cfg = ref.cfg
    joint_list, assoc = ref.paf_to_pose(heat, paf, cfg)
    humans_2d, visibility, conf_vec = ref.paf_to_human_list(joint_list, assoc)
    K = cfg.MODEL.NUM_KEYPOINTS
    humans_depth = []
    for i, human in enumerate(humans_2d):
        hd = np.ones(K) * -1
        for j, joint in enumerate(human):
            if visibility[i][j] > 0.5:
                hd[j] = ref.retrieve_depth_heat_weighted(
                    [int(joint[0] / cfg.MODEL.DOWNSAMPLE), int(joint[1] / cfg.MODEL.DOWNSAMPLE)],
                    posedepth[:, :, j], heat[:, :, j], radius=1)
        humans_depth.append(hd)
    for i, human in enumerate(humans_2d):
        human = np.array(human)
        human[np.where(visibility[i]), 0] = human[np.where(visibility[i]), 0] / 224 * cam.w_org
        human[np.where(visibility[i]), 1] = human[np.where(visibility[i]), 1] / 224 * cam.h_org
        humans_2d[i] = human
    humans_3d = []
    for i, human in enumerate(humans_2d):
        x3 = (human[:, 0] - cam.cx) * humans_depth[i] / cam.fx
        y3 = (human[:, 1] - cam.cy) * humans_depth[i] / cam.fy
        if cam.flip_y:
            y3 = -y3
        humans_3d.append(np.vstack([x3, y3, humans_depth[i]]).T.tolist())
        humans_2d[i] = human.tolist()
    return {"joint_list": np.asarray(joint_list, np.float64).reshape(-1, 5),
            "assoc": np.asarray(assoc, np.float64).reshape(-1, K + 2),
            "humans_2d": humans_2d, "humans_3d": humans_3d, "visibility": visibility,
            "conf": [[float(c) for c in cv] for cv in conf_vec]}
