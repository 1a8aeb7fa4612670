"""ORACLE (test infrastructure, NOT product code): CPU restatement of the reference decode + 3D lift.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (popnet_b200.decode) never does -- it calls the CUDA library and
fails loudly when that is missing.

What is restated (reference file:line, all under third_party_methods/):
  find_peaks                lib/utils/paf_to_pose.py:33-46
  NMS / refine              lib/utils/paf_to_pose.py:75-153
  find_connected_joints     lib/utils/paf_to_pose.py:156-264
  group_limbs_of_same_person lib/utils/paf_to_pose.py:267-351
  paf_to_pose               lib/utils/paf_to_pose.py:354-377
  paf_to_human_list         lib/utils/common.py:5-32
  retrieve_depth_heat_weighted lib/utils/common.py:272-293
  depth de-normalisation, 2D rescale, 3D back-projection
                            evaluate/evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:176-263

Third-party arithmetic on this path that is NOT vendored in the reference:
  * OpenCV ``cv2.resize(..., INTER_CUBIC)`` (pinned opencv-python==4.2.0.32, environment.yaml:325; 4.13.0
    in this image).  Restated below as ``bicubic_*``: Keys kernel A=-0.75, half-pixel centres, replicate
    border, horizontal pass ``((s0*a0+s1*a1)+s2*a2)+s3*a3`` then vertical pass
    ``s0*b0+(s1*b1+(s2*b2+s3*b3))``, all fp32 with separate multiply and add.  This is BIT-EXACT with
    OpenCV's own C++ code path (``cv2.ipp.setUseIPP(False)``; checked in tests/test_oracle_vs_golden.py
    through the committed fixtures).  With IPP enabled (the wheel's default) OpenCV dispatches to a
    closed-source Intel kernel whose rounding differs by <= 3.6e-7; fixtures for that variant are
    compared at the ">= 99.9 % of frames identical" bar instead.
  * SciPy ``maximum_filter`` with the 4-connected footprint, mode='reflect' (scipy=1.4.1 pinned).
  * NumPy reductions: pairwise sum order of ``np.sum`` / ``mean`` (restated in ``_np_sum_f32`` /
    ``_mean10``), and the float64 ``ndarray.dot`` of the 10x2 sample matrix with the unit direction.

PARITY PINNING: the reference ships no tests or golden vectors for this path.  The
oracle is pinned against outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py (which imports /root/reference read-only) and committed under
tests/golden/.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
f64 = np.float64

LIMBS_ITOP15 = ((8, 9), (9, 11), (11, 13), (8, 10), (10, 12), (12, 14), (8, 1), (1, 2), (2, 4), (4, 6),
                (1, 3), (3, 5), (5, 7), (1, 0))


# --------------------------------------------------------------------------------------------
# OpenCV INTER_CUBIC restatement
# --------------------------------------------------------------------------------------------
def cubic_coeffs(x) -> np.ndarray:
    """OpenCV interpolateCubic (imgproc/src/resize.cpp), evaluated in fp32."""
    x = f32(x); A = f32(-0.75); one = f32(1)
    c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
    c2 = ((A + f32(2)) * (one - x) - (A + f32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=f32)


def phase_table(scale: int = 8):
    """For destination index d = scale*q + r: source base offset (relative to q) and the 4 taps.
    Returns (ofs[scale] int, coef[scale,4] fp32): taps read source cells q+ofs-1 .. q+ofs+2."""
    ofs = np.zeros(scale, np.int64)
    coef = np.zeros((scale, 4), f32)
    for r in range(scale):
        fx = (r + 0.5) / scale - 0.5
        s = int(np.floor(fx))
        ofs[r] = s
        coef[r] = cubic_coeffs(fx - s)
    return ofs, coef


_OFS8, _COEF8 = phase_table(8)


def bicubic_upsample(src: np.ndarray, scale: int = 8) -> np.ndarray:
    """Full upsample of a 2-D fp32 array (== cv2.resize(src, None, fx=scale, fy=scale, INTER_CUBIC)
    on OpenCV's non-IPP path)."""
    src = np.ascontiguousarray(src, dtype=f32)
    H, W = src.shape
    ofs, coef = phase_table(scale)
    dx = np.arange(W * scale)
    qx, rx = dx // scale, dx % scale
    bx = qx + ofs[rx]
    ix = np.clip(bx[None, :] + np.arange(-1, 3)[:, None], 0, W - 1)            # [4, W*s]
    ax = coef[rx].T                                                              # [4, W*s]
    tmp = ((src[:, ix[0]] * ax[0] + src[:, ix[1]] * ax[1]) + src[:, ix[2]] * ax[2]) + src[:, ix[3]] * ax[3]
    tmp = tmp.astype(f32)
    dy = np.arange(H * scale)
    qy, ry = dy // scale, dy % scale
    by = qy + ofs[ry]
    iy = np.clip(by[None, :] + np.arange(-1, 3)[:, None], 0, H - 1)
    ay = coef[ry].T[:, :, None]
    out = tmp[iy[0]] * ay[0] + (tmp[iy[1]] * ay[1] + (tmp[iy[2]] * ay[2] + tmp[iy[3]] * ay[3]))
    return out.astype(f32)


def bicubic_sample(chan: np.ndarray, X: np.ndarray, Y: np.ndarray, scale: int = 8) -> np.ndarray:
    """Value of bicubic_upsample(chan)[Y, X] at integer points without materialising the upsample."""
    H, W = chan.shape
    ofs, coef = (_OFS8, _COEF8) if scale == 8 else phase_table(scale)
    X = np.asarray(X, np.int64); Y = np.asarray(Y, np.int64)
    rx, ry = X % scale, Y % scale
    bx, by = X // scale + ofs[rx], Y // scale + ofs[ry]
    ax, ay = coef[rx], coef[ry]                                                  # [n,4]
    rows = []
    for j in range(4):
        yy = np.clip(by - 1 + j, 0, H - 1)
        s = [chan[yy, np.clip(bx - 1 + i, 0, W - 1)] for i in range(4)]
        rows.append((((s[0] * ax[:, 0] + s[1] * ax[:, 1]) + s[2] * ax[:, 2]) + s[3] * ax[:, 3]).astype(f32))
    out = rows[0] * ay[:, 0] + (rows[1] * ay[:, 1] + (rows[2] * ay[:, 2] + rows[3] * ay[:, 3]))
    return out.astype(f32)


# --------------------------------------------------------------------------------------------
# D1 / D2: peaks
# --------------------------------------------------------------------------------------------
def find_peaks(thresh: float, img: np.ndarray) -> np.ndarray:
    """paf_to_pose.py:33-46: cell is a peak iff it equals the max over {itself, N, S, E, W} (neighbours
    outside the map ignored == scipy 'reflect') and is > thresh.  Row-major order, returned as [x, y]."""
    img = np.asarray(img)
    pad = np.pad(img, 1, mode="constant", constant_values=-np.inf)
    m = np.maximum.reduce([pad[1:-1, 1:-1], pad[:-2, 1:-1], pad[2:, 1:-1], pad[1:-1, :-2], pad[1:-1, 2:]])
    ys, xs = np.nonzero((m == img) & (img > thresh))
    return np.stack([xs, ys], 1) if len(xs) else np.zeros((0, 2), np.int64)


def nms(heat_chw: np.ndarray, num_keypoints: int, thresh: float, scale: int = 8, win: int = 2):
    """paf_to_pose.py:75-153 -> list over joint types of [n,4] float64 rows (X, Y, score, global id)."""
    out = []
    cnt = 0
    for j in range(num_keypoints):
        m = np.ascontiguousarray(heat_chw[j], dtype=f32)
        H, W = m.shape
        pk = find_peaks(thresh, m)
        rows = np.zeros((len(pk), 4), f64)
        for i, (x, y) in enumerate(pk):
            x0, y0 = max(0, x - win), max(0, y - win)
            x1, y1 = min(W - 1, x + win), min(H - 1, y + win)
            up = bicubic_upsample(m[y0:y1 + 1, x0:x1 + 1], scale)
            a = int(up.argmax())                      # first maximum, row-major
            ay, ax = divmod(a, up.shape[1])
            # (x+.5)*s-.5 + (ax - ((x-x0+.5)*s-.5))  ==  s*x0 + ax   (paf_to_pose.py:137-149)
            rows[i] = (scale * x0 + ax, scale * y0 + ay, up[ay, ax], cnt)
            cnt += 1
        out.append(rows)
    return out


# --------------------------------------------------------------------------------------------
# D4: limb scoring + greedy matching
# --------------------------------------------------------------------------------------------
def _line_points(a: int, b: int, n: int) -> np.ndarray:
    """round(linspace(a, b, n)) for integer endpoints; exact integer form floor((2((n-1)a+i(b-a))+(n-1))/(2(n-1)))
    (no half-way cases exist for n=10)."""
    i = np.arange(n)
    return np.round(np.linspace(f64(a), f64(b), num=n)).astype(np.int64) if n != 10 else \
        (2 * (9 * a + i * (b - a)) + 9) // 18


def _mean10(s: np.ndarray) -> f64:
    """np.mean of a length-n float64 vector: numpy pairwise sum (8-way unrolled for 8 <= n <= 128)."""
    n = len(s)
    if n < 8:
        r = f64(0.0)
        for v in s:
            r = r + v
    else:
        r8 = [s[k] for k in range(8)]
        i = 8
        while i + 8 <= n:
            for k in range(8):
                r8[k] = r8[k] + s[i + k]
            i += 8
        r = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]))
        while i < n:
            r = r + s[i]
            i += 1
    return r / f64(n)


def find_connected_joints(paf_chw: np.ndarray, peaks, limbs, *, thresh_paf: float, n_pts: int = 10,
                          scale: int = 8, paf_upsampled=None):
    """paf_to_pose.py:156-264.  ``paf_upsampled`` (optional [H*s, W*s, 2L]) switches from on-the-fly
    sampling to indexing a materialised upsample (used to validate the two are identical)."""
    Hup = paf_chw.shape[1] * scale
    connected = []
    for l, (ja, jb) in enumerate(limbs):
        src, dst = peaks[ja], peaks[jb]
        if len(src) == 0 or len(dst) == 0:
            connected.append(np.zeros((0, 5), f64) if False else [])
            continue
        cand = []
        px_map = np.ascontiguousarray(paf_chw[2 * l], f32)
        py_map = np.ascontiguousarray(paf_chw[2 * l + 1], f32)
        for i in range(len(src)):
            for j in range(len(dst)):
                dx = dst[j, 0] - src[i, 0]
                dy = dst[j, 1] - src[i, 1]
                dist = np.sqrt(dx * dx + dy * dy) + 1e-8
                ux, uy = dx / dist, dy / dist
                xs = _line_points(int(src[i, 0]), int(dst[j, 0]), n_pts)
                ys = _line_points(int(src[i, 1]), int(dst[j, 1]), n_pts)
                if paf_upsampled is None:
                    px = bicubic_sample(px_map, xs, ys, scale).astype(f64)
                    py = bicubic_sample(py_map, xs, ys, scale).astype(f64)
                else:
                    px = paf_upsampled[ys, xs, 2 * l].astype(f64)
                    py = paf_upsampled[ys, xs, 2 * l + 1].astype(f64)
                s = px * ux + py * uy
                score = _mean10(s) + min(0.5 * Hup / dist - 1, 0)
                if np.count_nonzero(s > thresh_paf) > 0.8 * n_pts and score > 0:
                    cand.append((i, j, score))
        # stable descending sort (Python sorted(reverse=True) keeps the original order of equal keys)
        order = sorted(range(len(cand)), key=lambda c: cand[c][2], reverse=True)
        conns = []
        used_i, used_j = set(), set()
        max_conn = min(len(src), len(dst))
        for c in order:
            i, j, s = cand[c]
            if i not in used_i and j not in used_j:
                conns.append((src[i, 3], dst[j, 3], s, i, j))
                used_i.add(i); used_j.add(j)
                if len(conns) >= max_conn:
                    break
        connected.append(np.array(conns, f64).reshape(-1, 5))
    return connected


# --------------------------------------------------------------------------------------------
# D5: person assembly
# --------------------------------------------------------------------------------------------
def group_limbs(connected, joint_list: np.ndarray, limbs, num_keypoints: int) -> np.ndarray:
    """paf_to_pose.py:267-351, including its quirks: overwrite of a different dst joint on a single
    match, merge adds p2+1, the non-disjoint 2-match path skips the 'dst differs' test, and 0 or >= 3
    matching persons open a new person."""
    K = num_keypoints
    persons = []
    for l, (ja, jb) in enumerate(limbs):
        for info in connected[l]:
            hits = [p for p, row in enumerate(persons) if row[ja] == info[0] or row[jb] == info[1]]
            if len(hits) == 1:
                row = persons[hits[0]]
                if row[jb] != info[1]:
                    row[jb] = info[1]
                    row[K + 1] += 1
                    row[K] += joint_list[int(info[1]), 2] + info[2]
            elif len(hits) == 2:
                r1, r2 = persons[hits[0]], persons[hits[1]]
                if not ((r1[:K] >= 0) & (r2[:K] >= 0)).any():
                    r1[:K] += r2[:K] + 1
                    r1[K:] += r2[K:]
                    r1[K] += info[2]
                    persons.pop(hits[1])
                else:
                    r1[jb] = info[1]
                    r1[K + 1] += 1
                    r1[K] += joint_list[int(info[1]), 2] + info[2]
            else:
                row = -np.ones(K + 2, f64)
                row[ja] = info[0]
                row[jb] = info[1]
                row[K + 1] = 2
                row[K] = (0 + joint_list[int(info[0]), 2] + joint_list[int(info[1]), 2]) + info[2]
                persons.append(row)
    keep = [r for r in persons if not (r[K + 1] < 3 or r[K] / r[K + 1] < 0.2)]
    return np.array(keep)


def paf_to_pose(heat_hwc: np.ndarray, paf_hwc: np.ndarray, *, num_keypoints=15, limbs=LIMBS_ITOP15,
                thresh_heat=0.1, thresh_paf=0.05, n_pts=10, scale=8):
    """paf_to_pose.py:354-377 on HWC maps, returns (joint_list [N,5] f64, person_to_joint_assoc [P,K+2] f64)."""
    heat = np.ascontiguousarray(np.transpose(heat_hwc, (2, 0, 1)), f32)
    paf = np.ascontiguousarray(np.transpose(paf_hwc, (2, 0, 1)), f32)
    return paf_to_pose_chw(heat, paf, num_keypoints=num_keypoints, limbs=limbs, thresh_heat=thresh_heat,
                           thresh_paf=thresh_paf, n_pts=n_pts, scale=scale)


def paf_to_pose_chw(heat, paf, *, num_keypoints=15, limbs=LIMBS_ITOP15, thresh_heat=0.1, thresh_paf=0.05,
                    n_pts=10, scale=8):
    peaks = nms(heat, num_keypoints, thresh_heat, scale)
    rows = [tuple(p) + (t,) for t, pk in enumerate(peaks) for p in pk]
    joint_list = np.array(rows, f64).reshape(-1, 5)
    connected = find_connected_joints(paf, peaks, limbs, thresh_paf=thresh_paf, n_pts=n_pts, scale=scale)
    assoc = group_limbs(connected, joint_list, limbs, num_keypoints)
    return joint_list, assoc


# --------------------------------------------------------------------------------------------
# D7 - D9: human list, depth lift, 3D
# --------------------------------------------------------------------------------------------
def paf_to_human_list(joint_list: np.ndarray, assoc: np.ndarray):
    """common.py:5-32."""
    humans, vis, conf = [], [], []
    for row in assoc:
        ids = row[:-2].astype(int)
        humans.append([[-1, -1] if i < 0 else joint_list[i, :2].tolist() for i in ids])
        conf.append([0 if i < 0 else float(joint_list[i, 2]) for i in ids])
        vis.append((ids >= 0).astype(int).tolist())
    return humans, vis, conf


def _np_sum_f32(a: np.ndarray) -> f32:
    """np.sum of a contiguous fp32 array of n <= 128 elements (numpy pairwise_sum, loops_utils.h): a plain loop below 8;
    else eight running sums over whole blocks of 8, combined as a tree, then the tail in order."""
    a = a.ravel()
    n = len(a)
    assert n <= 128
    if n < 8:
        r = f32(0.0)
        for v in a:
            r = f32(r + v)
        return r
    r8 = [f32(a[j]) for j in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r8[j] = f32(r8[j] + a[i + j])
        i += 8
    r = f32(f32(f32(r8[0] + r8[1]) + f32(r8[2] + r8[3])) + f32(f32(r8[4] + r8[5]) + f32(r8[6] + r8[7])))
    for v in a[i:]:
        r = f32(r + v)
    return r


def retrieve_depth_heat_weighted(center, depthmap: np.ndarray, heatmap: np.ndarray, radius: int = 1) -> f32:
    """common.py:272-293 (fp32 throughout; the in-place ``heatmap[heatmap<0]=0`` is applied to a copy)."""
    heatmap = np.where(heatmap < 0, f32(0), heatmap).astype(f32)
    gx, gy = depthmap.shape[1], depthmap.shape[0]
    x0 = min(max(int(center[0] - radius), 0), gx - 1); x1 = max(min(int(center[0] + radius), gx - 1), 0)
    y0 = min(max(int(center[1] - radius), 0), gy - 1); y1 = max(min(int(center[1] + radius), gy - 1), 0)
    w = (heatmap[y0:y1 + 1, x0:x1 + 1] + f32(0.000000001)).astype(f32)
    d = depthmap[y0:y1 + 1, x0:x1 + 1].astype(f32)
    return f32(_np_sum_f32((d * w).astype(f32)) / _np_sum_f32(w))


def retrieve_depth_weighted(center, depthmap: np.ndarray, radius: int = 1) -> f32:
    """common.py:251-269: np.mean of the clipped window (fp32 pairwise sum, fp32 division by the count)."""
    gx, gy = depthmap.shape[1], depthmap.shape[0]
    x0 = min(max(int(center[0] - radius), 0), gx - 1); x1 = max(min(int(center[0] + radius), gx - 1), 0)
    y0 = min(max(int(center[1] - radius), 0), gy - 1); y1 = max(min(int(center[1] + radius), gy - 1), 0)
    d = depthmap[y0:y1 + 1, x0:x1 + 1].astype(f32)
    return f32(_np_sum_f32(d) / f32(d.size))


def retrieve_depth_heat_max(center, depthmap: np.ndarray, heatmap: np.ndarray, radius: int = 1) -> f32:
    """common.py:296-318: depth at the first (row-major) maximum of max(heat, 0) in the clipped window."""
    heatmap = np.where(heatmap < 0, f32(0), heatmap).astype(f32)
    gx, gy = depthmap.shape[1], depthmap.shape[0]
    x0 = min(max(int(center[0] - radius), 0), gx - 1); x1 = max(min(int(center[0] + radius), gx - 1), 0)
    y0 = min(max(int(center[1] - radius), 0), gy - 1); y1 = max(min(int(center[1] + radius), gy - 1), 0)
    w = heatmap[y0:y1 + 1, x0:x1 + 1].ravel()
    d = depthmap[y0:y1 + 1, x0:x1 + 1].astype(f32).ravel()
    best = 0
    for i in range(1, len(w)):
        if w[i] > w[best]:
            best = i
    return f32(d[best])


def lift_frame(heat_chw, depth_chw, joint_list, assoc, *, scale=8, size=224, w_org=480, h_org=512,
               fx=504.1189880371094, fy=504.042724609375, cx=231.7421875, cy=320.62640380859375,
               depth_mean=3.0, depth_std=2.0, flip_y=False):
    """The per-frame body of evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:179-263.
    Returns (humans_2d [P][K][2], humans_3d [P][K][3], visibility [P][K], conf [P][K]) as float64 lists."""
    depth = (depth_chw.astype(f32) * f32(depth_std)).astype(f32)
    depth = (depth + f32(depth_mean)).astype(f32)
    humans, vis, conf = paf_to_human_list(joint_list, assoc)
    h2d, h3d = [], []
    for p, human in enumerate(humans):
        Z = np.ones(len(human), f64) * -1
        for j, (x, y) in enumerate(human):
            if vis[p][j] > 0.5:
                Z[j] = retrieve_depth_heat_weighted([int(x / scale), int(y / scale)], depth[j], heat_chw[j], 1)
        arr = np.array(human, f64)
        v = np.array(vis[p], bool)
        arr[v, 0] = arr[v, 0] / size * w_org
        arr[v, 1] = arr[v, 1] / size * h_org
        X3 = (arr[:, 0] - cx) * Z / fx
        Y3 = (arr[:, 1] - cy) * Z / fy
        if flip_y:
            Y3 = -Y3
        h2d.append(arr.tolist())
        h3d.append(np.vstack([X3, Y3, Z]).T.tolist())
    return h2d, h3d, vis, conf


def decode_frame(heat_chw, paf_chw, depth_chw, **kw):
    """paf_to_pose + lift on channel-major maps -- one frame of the hot path after the network."""
    lift_keys = ("w_org", "h_org", "fx", "fy", "cx", "cy", "depth_mean", "depth_std", "flip_y", "size")
    lkw = {k: kw.pop(k) for k in lift_keys if k in kw}
    joint_list, assoc = paf_to_pose_chw(heat_chw, paf_chw, **kw)
    h2d, h3d, vis, conf = lift_frame(heat_chw, depth_chw, joint_list, assoc, scale=kw.get("scale", 8), **lkw)
    return {"joint_list": joint_list, "assoc": assoc, "humans_2d": h2d, "humans_3d": h3d,
            "visibility": vis, "conf": conf}
