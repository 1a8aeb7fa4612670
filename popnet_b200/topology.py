"""Skeleton topology and camera constants of the MP-3DHP / ITOP depth-pose path.

These are the values the reference keeps as module globals:
  * joint names / limb list  -- util/util_functions.py:17-55 (identical copy in
    third_party_methods/lib/datasets/datasets_itop_rtpose.py:45-97, which is the one
    lib/utils/paf_to_pose.py:28-30 binds at import time)
  * MP-3DHP intrinsics, depth_mean/std/max -- util/util_functions.py:4,11-13
  * ITOP intrinsics -- third_party_methods/lib/datasets/datasets_itop_rtpose.py:32
They are data, not code: the decode kernels take them as a parameter block
(`DecodeParams`, include/popnet_b200.h) instead of baking them in.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

JOINT_NAMES: Tuple[str, ...] = (
    "head", "neck", "right_shoulder", "left_shoulder", "right_elbow", "left_elbow",
    "right_wrist", "left_wrist", "torso", "right_hip", "left_hip", "right_knee",
    "left_knee", "right_ankle", "left_ankle",
)

_LIMB_NAMES = (
    ("torso", "right_hip"), ("right_hip", "right_knee"), ("right_knee", "right_ankle"),
    ("torso", "left_hip"), ("left_hip", "left_knee"), ("left_knee", "left_ankle"),
    ("torso", "neck"), ("neck", "right_shoulder"), ("right_shoulder", "right_elbow"),
    ("right_elbow", "right_wrist"), ("neck", "left_shoulder"), ("left_shoulder", "left_elbow"),
    ("left_elbow", "left_wrist"), ("neck", "head"),
)

#: (src joint type, dst joint type) per limb; PAF channels (2i, 2i+1) = (x, y) of limb i.
LIMBS: Tuple[Tuple[int, int], ...] = tuple(
    (JOINT_NAMES.index(a), JOINT_NAMES.index(b)) for a, b in _LIMB_NAMES)

NUM_JOINTS = len(JOINT_NAMES)   # K = 15
NUM_LIMBS = len(LIMBS)          # L = 14
HEAD_ID, NECK_ID = 0, 1

MAX_JOINTS = 24                 # compile-time capacity of the CUDA decode (COCO's 18 fits)
MAX_LIMBS = 24

#: COCO body topology of the reference's other PAF users: 18 keypoints
#: (third_party_methods/lib/datasets/datasets_coco.py:40-64) and the 19 limbs of the native decoder
#: (third_party_methods/lib/pafprocess/pafprocess.h:21-24, COCOPAIRS).  The decode kernels take the topology as
#: data (DecodeParams.limbs), so this is a second parameter block, not a second code path.
COCO_JOINT_NAMES: Tuple[str, ...] = (
    "nose", "neck", "right_shoulder", "right_elbow", "right_wrist", "left_shoulder", "left_elbow", "left_wrist",
    "right_hip", "right_knee", "right_ankle", "left_hip", "left_knee", "left_ankle", "right_eye", "left_eye",
    "right_ear", "left_ear",
)
COCO_LIMBS: Tuple[Tuple[int, int], ...] = (
    (1, 2), (1, 5), (2, 3), (3, 4), (5, 6), (6, 7), (1, 8), (8, 9), (9, 10), (1, 11),
    (11, 12), (12, 13), (1, 0), (0, 14), (14, 16), (0, 15), (15, 17), (2, 16), (5, 17),
)


def get_keypoints() -> List[str]:
    """Same return value as util/util_functions.py:37-55."""
    return list(JOINT_NAMES)


def kp_connections(keypoints) -> List[List[int]]:
    """Same return value as util/util_functions.py:17-34."""
    return [[keypoints.index(a), keypoints.index(b)] for a, b in _LIMB_NAMES]


@dataclass(frozen=True)
class Camera:
    fx: float
    fy: float
    cx: float
    cy: float
    w_org: int
    h_org: int
    depth_mean: float = 3.0
    depth_std: float = 2.0
    depth_max: float = 6.0
    flip_y: bool = False


#: MP-3DHP Kinect camera (util/util_functions.py:4,11-13; 480x512 frames, evaluation_*_mpreal_ablation.py:60-63)
MP3DHP = Camera(fx=504.1189880371094, fy=504.042724609375, cx=231.7421875, cy=320.62640380859375,
                w_org=480, h_org=512, depth_mean=3.0, depth_std=2.0, depth_max=6.0)

#: ITOP camera (datasets_itop_rtpose.py:32-42): f = 1/0.0035, 320x240, Y negated (evaluation_rtpose_light3d_itop.py:203-208)
ITOP = Camera(fx=1.0 / 0.0035, fy=1.0 / 0.0035, cx=160.0, cy=120.0, w_org=320, h_org=240,
              depth_mean=3.0, depth_std=2.0, depth_max=5.0, flip_y=True)


@dataclass
class DecodeConfig:
    """Duck-typed stand-in for the five yacs fields the reference decode reads
    (SURVEY.md section 5): MODEL.NUM_KEYPOINTS, MODEL.DOWNSAMPLE, TEST.THRESH_HEATMAP,
    TEST.THRESH_PAF, TEST.NUM_INTERMED_PTS_BETWEEN_KEYPOINTS (lib/config/default.py:128-130)."""
    num_keypoints: int = NUM_JOINTS
    num_limbs: int = NUM_LIMBS
    downsample: int = 8
    thresh_heatmap: float = 0.1
    thresh_paf: float = 0.05
    num_intermed_pts: int = 10
    limbs: Tuple[Tuple[int, int], ...] = field(default_factory=lambda: LIMBS)

    @staticmethod
    def from_cfg(cfg) -> "DecodeConfig":
        """Accept the reference's yacs-style object (cfg.MODEL.*, cfg.TEST.*) or a DecodeConfig."""
        if isinstance(cfg, DecodeConfig):
            return cfg
        return DecodeConfig(
            num_keypoints=int(cfg.MODEL.NUM_KEYPOINTS),
            num_limbs=int(getattr(cfg.MODEL, "NUM_LIMBS", NUM_LIMBS)),
            downsample=int(cfg.MODEL.DOWNSAMPLE),
            thresh_heatmap=float(cfg.TEST.THRESH_HEATMAP),
            thresh_paf=float(cfg.TEST.THRESH_PAF),
            num_intermed_pts=int(cfg.TEST.NUM_INTERMED_PTS_BETWEEN_KEYPOINTS),
        )


def coco_config(**kw) -> DecodeConfig:
    """DecodeConfig of the COCO 18-keypoint / 19-limb topology (same thresholds as the depth path unless overridden)."""
    return DecodeConfig(num_keypoints=len(COCO_JOINT_NAMES), num_limbs=len(COCO_LIMBS), limbs=COCO_LIMBS, **kw)
