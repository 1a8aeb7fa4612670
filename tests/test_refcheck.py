"""Build-container-only checks against the UNMODIFIED reference imported from /root/reference (skipped wherever
that tree does not exist, e.g. on the GPU box).  They re-validate, on fresh random inputs that are not in the
committed fixtures, that the oracle + host logic reproduce the reference bit for bit -- including the edge cases the
reference's code paths define (missing joints, empty prediction lists, frames without GT, duplicate predictions,
degenerate single-joint boxes whose IoU is NaN)."""
import contextlib
import copy
import io
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import refshim  # noqa: E402

pytestmark = [pytest.mark.refcheck, pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")]


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


def _ragged(rng, N, K=15):
    pred2, pred3, conf, gt2, gt3 = [], [], [], [], []
    for _ in range(N):
        G = int(rng.integers(0, 5))
        P = int(rng.integers(0, 5)) if G else 0
        g2 = rng.uniform(0, 500, (G, K, 2)); g3 = rng.uniform(-2, 5, (G, K, 3))
        pick = rng.integers(0, max(G, 1), P)
        p2 = g2[pick] + rng.normal(0, 8, (P, K, 2)) if G else np.zeros((0, K, 2))
        p3 = g3[pick] + rng.normal(0, 0.08, (P, K, 3)) if G else np.zeros((0, K, 3))
        miss = rng.random((P, K)) < 0.2
        miss[:, 0] = False
        p2[miss] = -1.0
        if P >= 2 and rng.random() < 0.3:
            p2[0] = p2[1]; p3[0] = p3[1]
        if G and P and rng.random() < 0.15:                # degenerate single-joint GT and prediction: 0/0 IoU
            g2[0, 1:] = -1.0
            p2[0, 1:] = -1.0
            p2[0, 0] = g2[0, 0]
        c = rng.uniform(0.1, 1, (P, K)); c[p2[:, :, 0] == -1] = 0
        pred2.append(p2.tolist()); pred3.append(p3.tolist()); conf.append(c.tolist())
        gt2.append(g2.tolist()); gt3.append(g3.tolist())
    return pred2, pred3, conf, gt2, gt3


@pytest.fixture()
def on_oracle(oracle_lib, monkeypatch):
    from oracle.backend import OracleBackend
    from popnet_b200 import evaluate
    monkeypatch.setattr(evaluate, "_backend", OracleBackend())
    return evaluate


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_evaluator_edge_cases_match_reference(seed, on_oracle):
    ref = refshim.load()
    E = on_oracle
    rng = np.random.default_rng(seed)
    pred2, pred3, conf, gt2, gt3 = _ragged(rng, 150)
    names = ["j%d" % i for i in range(15)]
    for fref, fnew, args in (
        (ref.eval_pck.eval_human_dataset_2d, E.eval_human_dataset_2d, (pred2, gt2, 15, 20.0, 0.5)),
        (ref.eval_pck.eval_human_dataset_3d, E.eval_human_dataset_3d, (pred2, gt2, pred3, gt3, 15, 0.1, 0.5)),
    ):
        a = _quiet(fref, *args)
        b = _quiet(fnew, *args)
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x, np.float64), np.asarray(y, np.float64), equal_nan=True)
    # PCKh needs a non-degenerate head size only where it is used; run on the frames without degenerate GTs too
    a = _quiet(ref.eval_pck.eval_human_dataset_2d_PCKh, pred2, gt2, 0, 1, 15, 0.5, 0.5)
    b = _quiet(E.eval_human_dataset_2d_PCKh, pred2, gt2, 0, 1, 15, 0.5, 0.5)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x, np.float64), np.asarray(y, np.float64), equal_nan=True)
    a = _quiet(ref.eval_mAP.eval_ap_3D, pred3, copy.deepcopy(conf), gt3, [], names, 0.1)
    b = _quiet(E.eval_ap_3D, pred3, copy.deepcopy(conf), gt3, [], names, 0.1)
    assert np.array_equal(a, b, equal_nan=True)
    # head-rectangle variants: eval_pck.py:157-229, eval_mAP.py:210-269
    rects = [[[float(x), float(y), float(x + w), float(y + h)] for x, y, w, h in rng.uniform(5, 60, (len(g), 4))] for g in gt2]
    a = _quiet(ref.eval_pck.eval_human_dataset_2d_PCKh_rect, pred2, gt2, rects, 15, 0.5, 0.5)
    b = _quiet(E.eval_human_dataset_2d_PCKh_rect, pred2, gt2, rects, 15, 0.5, 0.5)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x, np.float64), np.asarray(y, np.float64), equal_nan=True)
    a = _quiet(ref.eval_mAP.eval_ap_mpii, pred2, copy.deepcopy(conf), gt2, [], rects, names, 0.5)
    b = _quiet(E.eval_ap_mpii, pred2, copy.deepcopy(conf), gt2, [], rects, names, 0.5)
    assert np.array_equal(a, b, equal_nan=True)


def test_decode_fresh_frames_match_reference(oracle_lib):
    """Fresh seeds (not in the fixtures), OpenCV's C++ resize path: bit-identical joints, scores, 2D / 3D poses."""
    cv2 = pytest.importorskip("cv2")
    import helpers
    from popnet_b200 import synth
    from popnet_b200.topology import MP3DHP
    ref = refshim.load()
    heat, paf, depth, _ = synth.map_batch(10, seed=777, persons=(1, 8), noise=0.015)
    out = oracle_lib.decode(heat, paf, depth, helpers.params_for("MP3DHP"))
    prev = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        for f in range(10):
            r = refshim.reference_decode_frame(ref, heat[f], paf[f], depth[f], MP3DHP)
            n = len(r["humans_2d"])
            g = {"k/joint_list": r["joint_list"], "k/assoc": r["assoc"],
                 "k/humans_2d": np.asarray(r["humans_2d"], np.float64).reshape(n, 15, 2),
                 "k/humans_3d": np.asarray(r["humans_3d"], np.float64).reshape(n, 15, 3),
                 "k/conf": np.asarray(r["conf"], np.float64).reshape(n, 15)}
            ok, why = helpers.compare_to_golden(out, f, g, "k/", exact=True)
            assert ok, (f, why)
    finally:
        cv2.ipp.setUseIPP(prev)


def test_decode_coco_topology_matches_reference(oracle_lib, monkeypatch):
    """the reference's paf_to_pose binds its skeleton as module globals (paf_to_pose.py:28-30);
    rebinding them to COCO's 18 keypoints / 19 limbs (pafprocess.h:21-24) must give what the oracle computes from the same
    topology passed as data."""
    cv2 = pytest.importorskip("cv2")
    import helpers
    from popnet_b200 import _abi, synth, topology
    ref = refshim.load()
    g = ref.paf_to_pose.__globals__
    limbs = [list(l) for l in topology.COCO_LIMBS]
    monkeypatch.setitem(g, "joint_to_limb_heatmap_relationship", limbs)
    monkeypatch.setitem(g, "paf_xy_coords_per_limb", np.arange(2 * len(limbs)).reshape(-1, 2))
    monkeypatch.setitem(g, "NUM_LIMBS", len(limbs))
    monkeypatch.setattr(ref.cfg.MODEL, "NUM_KEYPOINTS", 18)
    K = 18
    heat, paf, depth, _ = synth.map_batch(8, seed=99, persons=(1, 6), noise=0.01, limbs=topology.COCO_LIMBS,
                                          template=synth._TEMPLATE_COCO)
    out = oracle_lib.decode(heat, paf, depth, _abi.make_decode_params(topology.coco_config(), topology.MP3DHP))
    prev = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        for f in range(8):
            r = refshim.reference_decode_frame(ref, heat[f], paf[f], depth[f], topology.MP3DHP)
            n = len(r["humans_2d"])
            gold = {"k/joint_list": r["joint_list"], "k/assoc": r["assoc"],
                    "k/humans_2d": np.asarray(r["humans_2d"], np.float64).reshape(n, K, 2),
                    "k/humans_3d": np.asarray(r["humans_3d"], np.float64).reshape(n, K, 3),
                    "k/conf": np.asarray(r["conf"], np.float64).reshape(n, K)}
            ok, why = helpers.compare_to_golden(out, f, gold, "k/", K=K, exact=True)
            assert ok, (f, why)
    finally:
        cv2.ipp.setUseIPP(prev)


def test_depth_read_variants_match_reference():
    """oracle/decode_np.py restatements of retrieve_depth_weighted / retrieve_depth_heat_max vs lib/utils/common.py:251-318."""
    from oracle import decode_np
    ref = refshim.load()
    common = sys.modules[ref.paf_to_human_list.__module__]
    rng = np.random.default_rng(5)
    heat = (rng.random((28, 28), dtype=np.float32) - 0.2).astype(np.float32)
    heat[10:13, 4:7] = 0.5
    depth = (rng.random((28, 28), dtype=np.float32) * 4 + 1).astype(np.float32)
    for radius in (0, 1, 2, 3, 5):                     # windows of 1 .. 121 cells: every branch of NumPy's pairwise sum
        for y in range(28):
            for x in range(28):
                a = common.retrieve_depth_weighted((x, y), depth, radius=radius)
                b = decode_np.retrieve_depth_weighted((x, y), depth, radius)
                assert np.float32(a) == b and np.asarray(a).dtype == np.float32, (radius, x, y, a, b)
                a = common.retrieve_depth_heat_max((x, y), depth, heat.copy(), radius=radius)
                assert np.float32(a) == decode_np.retrieve_depth_heat_max((x, y), depth, heat, radius), (radius, x, y)
                a = common.retrieve_depth_heat_weighted((x, y), depth, heat.copy(), radius=radius)
                b = decode_np.retrieve_depth_heat_weighted((x, y), depth, heat, radius)
                assert np.float32(a) == b and np.asarray(a).dtype == np.float32, (radius, x, y, a, b)


@pytest.mark.parametrize("rows,cols", [(20, 28), (28, 17), (9, 40)], ids=lambda v: str(v))
def test_decode_non_square_maps_oracle_vs_reference(rows, cols, oracle_lib):
    """paf_to_pose (paf_to_pose.py:354-377) accepts any H x W: the C oracle on non-square crops of rendered maps is
    bit-identical to the reference (OpenCV C++ path) -- ids, coordinates, scores, associations."""
    import cv2
    import helpers
    from popnet_b200 import _abi, synth
    from popnet_b200.decode import records_to_reference
    from popnet_b200.topology import DecodeConfig, MP3DHP
    ref = refshim.load()
    heat, paf, depth, _ = synth.map_batch(6, seed=600 + rows, persons=(2, 6), noise=0.01, size=384)
    heat, paf, depth = (np.ascontiguousarray(t[:, :, 3:3 + rows, 5:5 + cols]) for t in (heat, paf, depth))
    params = _abi.make_decode_params(DecodeConfig(), MP3DHP, input_size=224, grid_hw=(rows, cols))
    out = oracle_lib.decode(heat, paf, depth, params)
    prev = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        for f in range(6):
            jl, assoc = ref.paf_to_pose(np.ascontiguousarray(heat[f].transpose(1, 2, 0)),
                                        np.ascontiguousarray(paf[f].transpose(1, 2, 0)), ref.cfg)
            ojl, oassoc = records_to_reference(out, f, 15)
            assert np.array_equal(np.asarray(jl, np.float64).reshape(-1, 5), ojl), f
            assert np.array_equal(np.asarray(assoc, np.float64).reshape(-1, 17), np.asarray(oassoc, np.float64).reshape(-1, 17)), f
    finally:
        cv2.ipp.setUseIPP(prev)
