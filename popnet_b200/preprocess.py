"""Eval-time input preprocessing on the device ("next" row 1 of SURVEY.md 8(f)).

Replaces, for a batch of raw metric depth frames,
  Cvt2ndarray + Resize(224) (cv2.INTER_LINEAR)       lib/datasets/data_augmentation_2d3d.py:70-89, 497-522
  clamp to [0, depth_max], (x - depth_mean) / depth_std   lib/datasets/datasets_kdh3d_rtpose_mpreal.py:CR229-246,
                                                          lib/datasets/datasets_itop_rtpose.py:213-223
with one bandwidth kernel (popnet_preprocess_depth).  No CPU fallback.
"""
from __future__ import annotations

from .topology import MP3DHP, Camera

_backend = None


def _get_backend():
    global _backend
    if _backend is None:
        from ._cuda_backend import CudaBackend
        _backend = CudaBackend()
    return _backend


def preprocess_depth(frames, camera: Camera = MP3DHP, size: int = 224):
    """frames [B, H, W] float32/float16 metres (NumPy array or CUDA tensor) -> CUDA tensor [B, 1, size, size]."""
    return _get_backend().preprocess_depth(frames, (size, size), camera.depth_max, camera.depth_mean, camera.depth_std)
