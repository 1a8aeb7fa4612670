"""GPU parity: CUDA PCK / mAP matching (through the C ABI) against the C oracle and the reference's numbers."""
import contextlib
import copy
import io

import numpy as np
import pytest

import helpers
from helpers import golden
from popnet_b200 import evaluate as E
from popnet_b200 import synth
from popnet_b200.topology import JOINT_NAMES

pytestmark = pytest.mark.gpu


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.fixture()
def eval_on_cuda(cuda_backend, monkeypatch):
    monkeypatch.setattr(E, "_backend", cuda_backend)
    return E


@pytest.mark.parametrize("tag,N,seed", [("small", 400, 7), ("c3", 4000, 0)])
def test_evaluator_bit_exact_vs_reference(tag, N, seed, eval_on_cuda):
    import warnings
    warnings.simplefilter("ignore")
    g = golden("eval_golden")
    ds = synth.eval_set(N, seed=seed)
    names = list(JOINT_NAMES)
    th2d = 0.02 * np.sqrt(480 ** 2 + 512 ** 2)
    a, k = E.eval_human_dataset_2d_PCKh(ds["pred2d"], ds["gt2d"], 0, 1, 15, 0.5, 0.5)
    assert np.array_equal(np.asarray(a), g[tag + "/pckh_avg"]) and np.array_equal(np.asarray(k), g[tag + "/pckh_kcp"])
    a, k = E.eval_human_dataset_2d(ds["pred2d"], ds["gt2d"], 15, th2d, 0.5)
    assert np.array_equal(np.asarray(a), g[tag + "/pck2d_avg"]) and np.array_equal(np.asarray(k), g[tag + "/pck2d_kcp"])
    a, k = E.eval_human_dataset_3d(ds["pred2d"], ds["gt2d"], ds["pred3d"], ds["gt3d"], 15, 0.1, 0.5)
    assert np.array_equal(np.asarray(a), g[tag + "/pck3d_avg"]) and np.array_equal(np.asarray(k), g[tag + "/pck3d_kcp"])
    ap2, c2 = _quiet(E.eval_ap_mpii_v2, ds["pred2d"], copy.deepcopy(ds["conf"]), ds["gt2d"], [], 0, 1, names, 0.5, _return_counts=True)
    ap3, c3 = _quiet(E.eval_ap_3D, ds["pred3d"], copy.deepcopy(ds["conf"]), ds["gt3d"], [], names, 0.1, _return_counts=True)
    # AP is a float64 from a sort + sums (device tail by default: stable tie order, different summation order than
    # NumPy's pairwise sum; the reference's own tie order is unspecified): tolerance, counters exact
    assert np.allclose(ap2, g[tag + "/ap2d"], rtol=0, atol=1e-9) and np.allclose(ap3, g[tag + "/ap3d"], rtol=0, atol=1e-9)
    E.AP_TAIL = "numpy"                  # the reference's own NumPy calls on the device's labels: bit-identical AP
    try:
        ap2n = _quiet(E.eval_ap_mpii_v2, ds["pred2d"], copy.deepcopy(ds["conf"]), ds["gt2d"], [], 0, 1, names, 0.5)
        ap3n = _quiet(E.eval_ap_3D, ds["pred3d"], copy.deepcopy(ds["conf"]), ds["gt3d"], [], names, 0.1)
    finally:
        E.AP_TAIL = "device"
    assert np.array_equal(ap2n, g[tag + "/ap2d"]) and np.array_equal(ap3n, g[tag + "/ap3d"])
    m = E.match_counts(ds["pred2d"], ds["gt2d"], num_joints=15, dist_th=th2d)
    assert np.array_equal(m["hit_cnt"], g[tag + "/hit_pck2d"]) and np.array_equal(m["valid_cnt"], g[tag + "/valid2d"])
    m3 = E.match_counts(ds["pred2d"], ds["gt2d"], pred3d=ds["pred3d"], gt3d=ds["gt3d"], num_joints=15, dist_th=0.1)
    assert np.array_equal(m3["hit_cnt"], g[tag + "/hit_pck3d"]) and np.array_equal(m3["valid_cnt"], g[tag + "/valid3d"])
    for dim, c in ((2, c2), (3, c3)):
        assert np.array_equal(c["n_pos"], g[tag + "/map%d_npos" % dim])
        assert np.array_equal(c["n_gt"], g[tag + "/map%d_ngt" % dim])
    if tag == "small":
        assert np.array_equal(m["dists"], g["small/dists2d"]) and np.array_equal(m3["dists"], g["small/dists3d"])
        assert np.array_equal(c2["labels"], g["small/map2_labels"]) and np.array_equal(c3["labels"], g["small/map3_labels"])


def _random_ragged(rng, N, K, maxh, dim3=True, p_missing=0.15, big=False):
    """Adversarial ragged set: empty frames, many humans, missing joints, duplicates, degenerate boxes."""
    pred2, pred3, conf, gt2, gt3, vis = [], [], [], [], [], []
    for _ in range(N):
        G = int(rng.integers(0, maxh + 1))
        P = int(rng.integers(0, maxh + 1)) if G > 0 else 0
        g2 = rng.uniform(0, 500, (G, K, 2)); g3 = rng.uniform(-2, 5, (G, K, 3))
        pick = rng.integers(0, max(G, 1), P)
        p2 = (g2[pick] + rng.normal(0, 8, (P, K, 2))) if G else np.zeros((0, K, 2))
        p3 = (g3[pick] + rng.normal(0, 0.08, (P, K, 3))) if G else np.zeros((0, K, 3))
        miss = rng.random((P, K)) < p_missing
        miss[:, 0] &= rng.random(P) < 0.5
        p2[miss] = -1.0
        if P and rng.random() < 0.1:
            p2[0] = p2[min(1, P - 1)]          # duplicate prediction: tie in the IoU arg-max
        c = rng.uniform(0, 1, (P, K)); c[miss] = 0
        v = (rng.random((G, K)) > 0.1).astype(float)
        if G:
            v[:, 0] = 1
        pred2.append(p2.tolist()); pred3.append(p3.tolist()); conf.append(c.tolist())
        gt2.append(g2.tolist()); gt3.append(g3.tolist()); vis.append(v.tolist())
    return pred2, pred3, conf, gt2, gt3, vis


@pytest.mark.parametrize("maxh,N", [(3, 300), (12, 200), (70, 12)], ids=["small", "crowd", "over-cache"])
def test_evaluator_random_ragged_vs_oracle(maxh, N, cuda_backend, oracle_lib):
    """CUDA vs C oracle on the raw kernels' outputs, with visibility masks and per-GT thresholds."""
    from oracle.backend import OracleBackend
    ob = OracleBackend()
    rng = np.random.default_rng(maxh)
    K = 15
    pred2, pred3, conf, gt2, gt3, vis = _random_ragged(rng, N, K, maxh)
    p2, poff = E.pack_humans(pred2, K, 2); g2, goff = E.pack_humans(gt2, K, 2)
    p3, _ = E.pack_humans(pred3, K, 3); g3, _ = E.pack_humans(gt3, K, 3)
    visf = np.ascontiguousarray((np.asarray([r for fr in vis for r in fr]).reshape(-1, K) > 0).astype(np.uint8))
    th = rng.uniform(5, 40, g2.shape[0])
    for use3d in (False, True):
        arrs = {"pred2d": p2, "pred_off": poff, "gt2d": g2, "gt_off": goff, "gt_vis": visf, "gt_thresh": th}
        if use3d:
            arrs.update(pred3d=p3, gt3d=g3)
            arrs.pop("gt_thresh")
        a = cuda_backend.pck(arrs, dist_th=0.1, iou_th=0.5, K=K)
        b = ob.pck(arrs, dist_th=0.1, iou_th=0.5, K=K)
        for k in b:
            assert np.array_equal(a[k], b[k]), k
    ref = rng.uniform(20, 80, g2.shape[0])
    for D, pp, gg, t in ((2, p2, g2, 0.5), (3, p3, g3, 0.1)):
        arrs = {"pred": pp, "pred_off": poff, "gt": gg, "gt_off": goff, "gt_vis": visf,
                "ref_dist": ref if D == 2 else np.ones_like(ref)}
        # frames with predictions but no GT cannot occur here (P = 0 whenever G = 0)
        a = cuda_backend.map_assign(arrs, thresh=t, K=K, D=D)
        b = ob.map_assign(arrs, thresh=t, K=K, D=D)
        for k in b:
            assert np.array_equal(a[k], b[k]), k


def test_evaluator_properties_full_size(eval_on_cuda):
    """Size-independent properties at C3's full size: permuting frames leaves every counter unchanged;
    concatenating two shards adds their counters (the all-reduce contract of the multi-GPU path)."""
    ds = synth.eval_set(4000, seed=0)
    th2d = 0.02 * np.sqrt(480 ** 2 + 512 ** 2)
    full = E.match_counts(ds["pred2d"], ds["gt2d"], pred3d=ds["pred3d"], gt3d=ds["gt3d"], num_joints=15, dist_th=0.1)
    perm = np.random.default_rng(1).permutation(4000)
    pm = E.match_counts([ds["pred2d"][i] for i in perm], [ds["gt2d"][i] for i in perm],
                        pred3d=[ds["pred3d"][i] for i in perm], gt3d=[ds["gt3d"][i] for i in perm],
                        num_joints=15, dist_th=0.1)
    assert np.array_equal(full["hit_cnt"], pm["hit_cnt"]) and np.array_equal(full["valid_cnt"], pm["valid_cnt"])
    halves = [E.match_counts(ds["pred2d"][s], ds["gt2d"][s], pred3d=ds["pred3d"][s], gt3d=ds["gt3d"][s],
                             num_joints=15, dist_th=0.1) for s in (slice(0, 1500), slice(1500, 4000))]
    assert np.array_equal(full["hit_cnt"], halves[0]["hit_cnt"] + halves[1]["hit_cnt"])
    assert full["samples_cnt"] == halves[0]["samples_cnt"] + halves[1]["samples_cnt"]
    # perfect predictions: every visible joint is a hit
    perfect = E.match_counts(ds["gt2d"], ds["gt2d"], num_joints=15, dist_th=th2d)
    assert (perfect["hit_cnt"] == perfect["samples_cnt"]).all()


def test_evaluator_edge_cases_gpu(eval_on_cuda):
    K = 15
    base = np.stack([np.linspace(100, 200, K), np.linspace(50, 400, K)], 1)
    allmiss = (-np.ones((K, 2))).tolist()
    one = -np.ones((K, 2)); one[3] = [120.0, 80.0]
    gt = [[base.tolist()], [], [base.tolist(), (base + 150).tolist()], [base.tolist()], [one.tolist()]]
    pred = [[], [], [(base + 3).tolist(), (base + 149).tolist()], [(base + 1).tolist(), allmiss], [one.tolist()]]
    m = E.match_counts(pred, gt, num_joints=K, dist_th=10.0)
    assert m["matched_pred"].tolist() == [-1, 0, 1, -1, 0]
    assert (m["dists"][3] == -1).all() and m["dists"][4][3] == 0.0
    with pytest.raises(IndexError):
        E.eval_human_dataset_2d([[base.tolist()]], [[allmiss]], K, 10.0, 0.5)


def test_head_rectangle_variants_cuda_vs_oracle(cuda_backend, monkeypatch):
    """eval_human_dataset_2d_PCKh_rect / eval_ap_mpii (eval_pck.py:157-229, eval_mAP.py:210-269): same numbers from the
    CUDA backend and the C oracle (tests/test_refcheck.py pins the oracle path to the reference)."""
    from oracle.backend import OracleBackend
    ds = synth.eval_set(600, seed=21)
    rng = np.random.default_rng(4)
    rects = [[[float(x), float(y), float(x + w), float(y + h)] for x, y, w, h in rng.uniform(5, 60, (len(g), 4))] for g in ds["gt2d"]]
    names = list(JOINT_NAMES)
    res = []
    for be in (cuda_backend, OracleBackend()):
        monkeypatch.setattr(E, "_backend", be)
        a = _quiet(E.eval_human_dataset_2d_PCKh_rect, ds["pred2d"], ds["gt2d"], rects, 15, 0.5, 0.5)
        b = _quiet(E.eval_ap_mpii, ds["pred2d"], copy.deepcopy(ds["conf"]), ds["gt2d"], [], rects, names, 0.5)
        res.append((np.asarray(a[0]), np.asarray(a[1]), np.asarray(b)))
    for x, y in zip(res[0][:2], res[1][:2]):           # PCKh values: identical counters -> identical numbers
        assert np.array_equal(x, y, equal_nan=True)
    # AP: the CUDA backend runs the tail on the device (the default), the oracle backend runs the NumPy tail: float64
    # sums in a different order -> 1e-9 on the 0..100 scale (as in test_ap_tail_on_device)
    assert np.allclose(res[0][2], res[1][2], rtol=0, atol=1e-9, equal_nan=True), np.abs(res[0][2] - res[1][2]).max()
    assert res[0][2][-1] > 10.0          # a meaningful AP, not an all-zero agreement


@pytest.mark.parametrize("N,seed", [(400, 7), (4000, 0)], ids=["small", "c3"])
def test_ap_tail_on_device(N, seed, cuda_backend, monkeypatch):
    """popnet_eval_ap (sort + precision/recall scan + VOC envelope on the device) against the NumPy tail that is
    bit-identical to util/eval_mAP.py:160-207; float64 sums in a different order -> 1e-9 on the 0..100 scale."""
    monkeypatch.setattr(E, "_backend", cuda_backend)
    ds = synth.eval_set(N, seed=seed)
    names = list(JOINT_NAMES)
    res = {}
    for tail in ("numpy", "device"):
        monkeypatch.setattr(E, "AP_TAIL", tail)
        res[tail] = (_quiet(E.eval_ap_3D, ds["pred3d"], copy.deepcopy(ds["conf"]), ds["gt3d"], [], names, 0.1),
                     _quiet(E.eval_ap_mpii_v2, ds["pred2d"], copy.deepcopy(ds["conf"]), ds["gt2d"], [], 0, 1, names, 0.5))
    for a, b in zip(res["numpy"], res["device"]):
        assert a.shape == b.shape == (16,)
        assert np.allclose(a, b, rtol=0, atol=1e-9), np.abs(a - b).max()
        assert a[-1] > 10


def test_ap_tail_sort_sizes_and_ties(cuda_backend):
    """Sizes around the shared-memory / global bitonic boundaries, duplicate scores with equal labels (order-free),
    zero-score missing joints, joints without any GT."""
    rng = np.random.default_rng(11)
    for SP in (0, 1, 5, 2047, 2048, 2049, 5000, 70000):
        K = 15
        conf = rng.uniform(0.05, 1.0, (SP, K))
        labels = (rng.random((SP, K)) < 0.6).astype(np.uint8)
        if SP:
            dup = rng.integers(0, SP, SP // 3)
            conf[dup] = np.round(conf[dup], 2)                  # many ties ...
            labels[dup] = 1                                     # ... whose labels agree, so any tie order gives one AP
            same = np.isin(conf, np.round(conf[dup], 2))
            labels[same] = 1
            miss = rng.random((SP, K)) < 0.1
            conf[miss] = 0.0
            labels[miss] = 0
        n_gt = labels.sum(0).astype(np.int64) + rng.integers(0, 50, K)
        n_gt[3] = 0 if SP == 0 else n_gt[3]
        labels[:, 7] = 0                                        # a joint that is never right
        got = cuda_backend.ap_tail(conf, labels, n_gt)
        want = np.zeros(K + 1)
        with np.errstate(all="ignore"):
            for j in range(K):
                pr, rc = E._get_rpc(conf[:, j], labels[:, j].astype(np.int64), np.float64(n_gt[j]))
                want[j] = E._voc_ap(rc, pr) * 100
        want[-1] = np.mean(want[:-1])
        assert np.allclose(got, want, rtol=0, atol=1e-9, equal_nan=True), (SP, np.abs(got - want).max())
