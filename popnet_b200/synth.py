"""Seeded synthetic workloads for the depth-pose path (no dataset or checkpoint ships with the
reference: labels/, predictions/ and trained_model/ are only in the 800 GB torrent, README.md:43-45).

Three generators, all deterministic in their seed:

* ``render_maps``   -- GT-style network outputs (joint heat-maps, part-affinity fields, per-joint depth
  maps) drawn from random skeletons.  The formulas are the ones the reference uses to render its
  training targets (Gaussian sigma 7 px masked at exponent 4.6052 and clamped at 1.0,
  lib/datasets/heatmap.py:20-36; unit vectors within one grid cell of the limb segment averaged over
  overlaps, lib/datasets/paf.py:18-69; constant-depth patches around the joint cell, min-composited,
  lib/datasets/posemap.py:83-106), re-implemented here in vectorised NumPy.  The clamp produces
  plateaus, which is what exercises the NMS tie rule.
* ``depth_frames``  -- metric depth images (background plane + capsule bodies, z-buffer composited,
  4 % zero holes) normalised like datasets_kdh3d_rtpose_mpreal.py:CR229-246.
* ``eval_set``      -- a multi-person prediction / ground-truth set for the PCK / mAP evaluator
  (SURVEY.md section 8(d), config C3).
"""
from __future__ import annotations

import numpy as np

from .topology import LIMBS, NUM_JOINTS, NUM_LIMBS, MP3DHP, Camera

# Standing-person template: x in [-0.5, 0.5] (fraction of body height), y in [0, 1] from head top to ankle.
_TEMPLATE = np.array([
    [0.00, 0.04], [0.00, 0.17], [-0.17, 0.21], [0.17, 0.21], [-0.24, 0.39], [0.24, 0.39],
    [-0.27, 0.55], [0.27, 0.55], [0.00, 0.42], [-0.09, 0.55], [0.09, 0.55], [-0.11, 0.75],
    [0.11, 0.75], [-0.12, 0.96], [0.12, 0.96]], dtype=np.float64)


#: COCO 18-keypoint template (topology.COCO_JOINT_NAMES order), same normalisation
_TEMPLATE_COCO = np.array([
    [0.00, 0.06], [0.00, 0.17], [-0.17, 0.21], [-0.24, 0.39], [-0.27, 0.55], [0.17, 0.21], [0.24, 0.39], [0.27, 0.55],
    [-0.09, 0.55], [-0.11, 0.75], [-0.12, 0.96], [0.09, 0.55], [0.11, 0.75], [0.12, 0.96],
    [-0.03, 0.03], [0.03, 0.03], [-0.07, 0.05], [0.07, 0.05]], dtype=np.float64)


def random_skeletons(rng: np.random.Generator, n_persons: int, size: int = 224,
                     height_range=(70.0, 190.0), jitter: float = 0.025, template: np.ndarray | None = None):
    """Return (joints2d [P,K,2] float64 in network-input pixels, z [P] metres); K = 15 unless `template` is given."""
    _T = _TEMPLATE if template is None else template
    out = np.zeros((n_persons, _T.shape[0], 2), np.float64)
    zs = np.zeros(n_persons, np.float64)
    for p in range(n_persons):
        h = rng.uniform(*height_range)
        cx = rng.uniform(0.12 * size, 0.88 * size)
        top = rng.uniform(-0.05 * size, size - 0.75 * h)
        lean = rng.uniform(-0.25, 0.25)
        pts = _T.copy()
        pts += rng.normal(0.0, jitter, pts.shape)
        pts[:, 0] += lean * (pts[:, 1] - 0.5)
        out[p, :, 0] = cx + pts[:, 0] * h
        out[p, :, 1] = top + pts[:, 1] * h
        zs[p] = rng.uniform(1.5, 4.5)
    return out, zs


def render_maps(joints2d: np.ndarray, z: np.ndarray, *, size: int = 224, stride: int = 8,
                sigma: float = 7.0, cam: Camera = MP3DHP, noise: float = 0.0,
                rng: np.random.Generator | None = None, limbs=None):
    """Render one frame's maps in the network's output layout (channel-major, fp32).

    Returns heat [K+1, g, g], paf [2L, g, g], depth [K, g, g] with g = size // stride; depth is
    normalised ``(z - depth_mean) / depth_std`` like the network's third head.
    """
    g = size // stride
    P = joints2d.shape[0]
    LIMBS = globals()["LIMBS"] if limbs is None else tuple(limbs)      # topology: default = the depth path's 15 / 14
    NUM_JOINTS = joints2d.shape[1] if limbs is not None else globals()["NUM_JOINTS"]
    NUM_LIMBS = len(LIMBS)
    centres = np.arange(g, dtype=np.float64) * stride + (stride / 2.0 - 0.5)
    xx, yy = np.meshgrid(centres, centres)
    heat = np.zeros((NUM_JOINTS + 1, g, g), np.float64)
    for p in range(P):
        for k in range(NUM_JOINTS):
            x, y = joints2d[p, k]
            if not (0 <= x < size and 0 <= y < size):
                continue
            e = ((xx - x) ** 2 + (yy - y) ** 2) / 2.0 / sigma / sigma
            heat[k] += np.where(e <= 4.6052, np.exp(-e), 0.0)
            np.minimum(heat[k], 1.0, out=heat[k])
    heat[NUM_JOINTS] = 1.0 - heat[:NUM_JOINTS].max(axis=0)

    paf = np.zeros((2 * NUM_LIMBS, g, g), np.float64)
    gi = np.arange(g, dtype=np.float64)
    gx, gy = np.meshgrid(gi, gi)
    for l, (a, b) in enumerate(LIMBS):
        cnt = np.zeros((g, g), np.float64)
        for p in range(P):
            A = joints2d[p, a] / stride
            B = joints2d[p, b] / stride
            if not (np.all(A >= -1) and np.all(B >= -1) and np.all(A <= g) and np.all(B <= g)):
                continue
            v = B - A
            n = np.hypot(v[0], v[1])
            if n == 0.0:
                continue
            u = v / n
            box = ((gx >= round(min(A[0], B[0]) - 1)) & (gx <= round(max(A[0], B[0]) + 1)) &
                   (gy >= round(min(A[1], B[1]) - 1)) & (gy <= round(max(A[1], B[1]) + 1)))
            m = box & (np.abs((gx - A[0]) * u[1] - (gy - A[1]) * u[0]) < 1.0)
            paf[2 * l] = (paf[2 * l] * cnt + m * u[0])
            paf[2 * l + 1] = (paf[2 * l + 1] * cnt + m * u[1])
            cnt += m
            d = np.where(cnt == 0, 1.0, cnt)
            paf[2 * l] /= d
            paf[2 * l + 1] /= d

    depth = np.full((NUM_JOINTS, g, g), cam.depth_max, np.float64)
    for p in range(P):
        for k in range(NUM_JOINTS):
            x, y = joints2d[p, k]
            if not (0 <= x < size and 0 <= y < size):
                continue
            cx, cy = int(x / stride), int(y / stride)
            x0, x1 = max(cx - 2, 0), min(cx + 2, g - 1)
            y0, y1 = max(cy - 2, 0), min(cy + 2, g - 1)
            patch = depth[k, y0:y1 + 1, x0:x1 + 1]
            np.minimum(patch, z[p] + 0.01 * (k % 3), out=patch)
    depth = (depth - cam.depth_mean) / cam.depth_std

    heat = heat.astype(np.float32)
    paf = paf.astype(np.float32)
    depth = depth.astype(np.float32)
    if noise > 0.0:
        assert rng is not None
        heat = np.clip(heat + rng.normal(0, noise, heat.shape).astype(np.float32), 0.0, 1.0).astype(np.float32)
        paf = (paf + rng.normal(0, noise, paf.shape).astype(np.float32)).astype(np.float32)
        depth = (depth + rng.normal(0, noise * 0.5, depth.shape).astype(np.float32)).astype(np.float32)
    return heat, paf, depth


def map_batch(batch: int, *, seed: int = 1234, persons=(1, 6), noise: float = 0.01, size: int = 224,
              stride: int = 8, limbs=None, template: np.ndarray | None = None):
    """C2/C4/C5 decode inputs: ``batch`` frames, persons ~ U{lo..hi} per frame, frame f seeded with
    ``default_rng(seed + f)``.  Returns (heat [B,16,g,g], paf [B,28,g,g], depth [B,15,g,g], skeletons)."""
    g = size // stride
    K = NUM_JOINTS if template is None else template.shape[0]
    L = NUM_LIMBS if limbs is None else len(limbs)
    heat = np.zeros((batch, K + 1, g, g), np.float32)
    paf = np.zeros((batch, 2 * L, g, g), np.float32)
    depth = np.zeros((batch, K, g, g), np.float32)
    skels = []
    for f in range(batch):
        rng = np.random.default_rng(seed + f)
        n = int(rng.integers(persons[0], persons[1] + 1))
        j2d, z = random_skeletons(rng, n, size, template=template)
        # every third frame keeps the clean (plateau-rich) rendering; the rest get sensor-like noise
        nz = 0.0 if f % 3 == 0 else noise
        heat[f], paf[f], depth[f] = render_maps(j2d, z, size=size, stride=stride, noise=nz, rng=rng, limbs=limbs)
        skels.append((j2d, z))
    return heat, paf, depth, skels


def depth_frames(batch: int, *, seed: int = 1234, persons=(1, 6), size: int = 224, cam: Camera = MP3DHP):
    """Normalised synthetic depth frames ``[B,1,size,size]`` fp32."""
    out = np.zeros((batch, 1, size, size), np.float32)
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float64)
    for f in range(batch):
        rng = np.random.default_rng(seed + f)
        n = int(rng.integers(persons[0], persons[1] + 1))
        j2d, z = random_skeletons(rng, n, size)
        img = np.clip(rng.normal(3.5, 0.3) + 0.002 * (ys - size / 2) + rng.normal(0, 0.01, (size, size)), 0, cam.depth_max)
        for p in range(n):
            r = 0.035 * (j2d[p, 13, 1] - j2d[p, 0, 1]) + 2.0
            for a, b in LIMBS:
                A, B = j2d[p, a], j2d[p, b]
                v = B - A
                L2 = max(float(v @ v), 1e-9)
                t = np.clip(((xs - A[0]) * v[0] + (ys - A[1]) * v[1]) / L2, 0.0, 1.0)
                d2 = (xs - A[0] - t * v[0]) ** 2 + (ys - A[1] - t * v[1]) ** 2
                img = np.where((d2 < r * r) & (z[p] < img), z[p], img)
        holes = rng.random((size, size)) < 0.04
        img = np.where(holes, 0.0, img)
        img = np.clip(img, 0.0, cam.depth_max)
        out[f, 0] = ((img - cam.depth_mean) / cam.depth_std).astype(np.float32)
    return out


def eval_set(n_frames: int = 4000, *, seed: int = 0, max_gt: int = 6, cam: Camera = MP3DHP):
    """C3: ragged prediction / GT lists in the reference's JSON layout.

    Returns dict with ``pred2d, pred3d, conf, gt2d, gt3d`` -- lists (frames) of lists (humans) of
    [K][2|3] lists, ``conf`` [K] per human; missing predicted joints are ``[-1, -1]`` with conf 0.
    """
    rng = np.random.default_rng(seed)
    pred2d, pred3d, conf, gt2d, gt3d = [], [], [], [], []
    for _ in range(n_frames):
        G = int(rng.integers(1, max_gt + 1))
        f_gt2, f_gt3, f_p2, f_p3, f_c = [], [], [], [], []
        for _g in range(G):
            h = rng.uniform(150.0, 400.0)
            cx = rng.uniform(0.15 * cam.w_org, 0.85 * cam.w_org)
            top = rng.uniform(0.0, max(cam.h_org - h, 1.0))
            pts = _TEMPLATE + rng.normal(0, 0.02, _TEMPLATE.shape)
            x = cx + pts[:, 0] * h
            y = top + pts[:, 1] * h
            Z = rng.uniform(1.5, 4.5) + rng.normal(0, 0.05, NUM_JOINTS)
            X3 = (x - cam.cx) * Z / cam.fx
            Y3 = (y - cam.cy) * Z / cam.fy
            f_gt2.append(np.stack([x, y], 1).tolist())
            f_gt3.append(np.stack([X3, Y3, Z], 1).tolist())
            if rng.random() < 0.9:
                px = x + rng.normal(0, 6.0, NUM_JOINTS)
                py = y + rng.normal(0, 6.0, NUM_JOINTS)
                pZ = Z + rng.normal(0, 0.06, NUM_JOINTS)
                c = rng.uniform(0.2, 1.0, NUM_JOINTS)
                miss = rng.random(NUM_JOINTS) < 0.1
                p2 = np.stack([px, py], 1)
                p3 = np.stack([(px - cam.cx) * pZ / cam.fx, (py - cam.cy) * pZ / cam.fy, pZ], 1)
                p2[miss] = -1.0
                p3[miss] = np.stack([(-1 - cam.cx) * -1 / cam.fx, (-1 - cam.cy) * -1 / cam.fy, -1.0])
                c[miss] = 0.0
                f_p2.append(p2.tolist()); f_p3.append(p3.tolist()); f_c.append(c.tolist())
        if rng.random() < 0.1:   # one false positive
            h = rng.uniform(150.0, 400.0)
            cx = rng.uniform(0.15 * cam.w_org, 0.85 * cam.w_org)
            top = rng.uniform(0.0, max(cam.h_org - h, 1.0))
            px = cx + _TEMPLATE[:, 0] * h
            py = top + _TEMPLATE[:, 1] * h
            pZ = np.full(NUM_JOINTS, rng.uniform(1.5, 4.5))
            f_p2.append(np.stack([px, py], 1).tolist())
            f_p3.append(np.stack([(px - cam.cx) * pZ / cam.fx, (py - cam.cy) * pZ / cam.fy, pZ], 1).tolist())
            f_c.append(rng.uniform(0.2, 1.0, NUM_JOINTS).tolist())
        gt2d.append(f_gt2); gt3d.append(f_gt3); pred2d.append(f_p2); pred3d.append(f_p3); conf.append(f_c)
    return {"pred2d": pred2d, "pred3d": pred3d, "conf": conf, "gt2d": gt2d, "gt3d": gt3d}
