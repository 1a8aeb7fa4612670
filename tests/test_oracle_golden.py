"""CPU suite, part 1: the oracle against the reference's own outputs (committed golden fixtures).

The fixtures were produced by tests/golden/make_golden.py, which runs the unmodified reference from
/root/reference; nothing here reads /root/reference.
"""
import copy
import io
import contextlib

import numpy as np
import pytest

import helpers
from helpers import DECODE_CASES, golden
from popnet_b200 import evaluate as E
from popnet_b200 import synth
from popnet_b200.topology import JOINT_NAMES


@pytest.fixture(scope="module")
def oracle_backend(oracle_lib):
    from oracle.backend import OracleBackend
    return OracleBackend()


@pytest.mark.parametrize("case", DECODE_CASES, ids=[c[0] for c in DECODE_CASES])
def test_c_oracle_decode_matches_reference_bitwise(case, oracle_lib):
    """Reference run with OpenCV's own C++ resize (IPP off): every id, coordinate, score, 2D/3D joint
    and confidence must be bit-identical."""
    g = golden("decode_golden")
    heat, paf, depth = helpers.decode_case_inputs(case)
    out = oracle_lib.decode(heat, paf, depth, helpers.params_for(case[5]))
    assert out["flags"].sum() == 0
    for f in range(heat.shape[0]):
        ok, why = helpers.compare_to_golden(out, f, g, "%s/native/%d/" % (case[0], f), exact=True)
        assert ok, "frame %d: %s" % (f, why)


@pytest.mark.parametrize("case", DECODE_CASES, ids=[c[0] for c in DECODE_CASES])
def test_c_oracle_decode_matches_reference_ipp(case, oracle_lib):
    """Reference run with the wheel's default (closed-source IPP) resize: scores differ by <= 3.6e-7,
    assembled joints must be identical on >= 99.9 % of frames (here: all of them)."""
    g = golden("decode_golden")
    heat, paf, depth = helpers.decode_case_inputs(case)
    out = oracle_lib.decode(heat, paf, depth, helpers.params_for(case[5]))
    bad = [f for f in range(heat.shape[0])
           if not helpers.compare_to_golden(out, f, g, "%s/ipp/%d/" % (case[0], f), exact=False)[0]]
    assert len(bad) <= 0.001 * heat.shape[0], bad


def test_numpy_oracle_agrees_with_c_oracle(oracle_lib):
    from oracle import decode_np
    heat, paf, depth = helpers.decode_case_inputs(DECODE_CASES[0])
    out = oracle_lib.decode(heat[:6], paf[:6], depth[:6], helpers.params_for("MP3DHP"))
    from popnet_b200.decode import records_to_reference
    for f in range(6):
        r = decode_np.decode_frame(heat[f], paf[f], depth[f])
        jl, assoc = records_to_reference(out, f, 15)
        assert np.array_equal(r["joint_list"], jl)
        a = np.asarray(r["assoc"]).reshape(-1, 17)
        assoc = np.asarray(assoc).reshape(-1, 17)
        assert np.array_equal(a[:, :15], assoc[:, :15]) and np.array_equal(a[:, 16], assoc[:, 16])
        # the NumPy restatement does not model the BLAS FMA pattern of ndarray.dot: scores agree to 1e-12
        assert np.allclose(a[:, 15], assoc[:, 15], rtol=0, atol=1e-12)
        n = len(a)
        assert np.array_equal(np.asarray(r["humans_2d"]).reshape(n, 15, 2), out["pose2d"][f, :n])
        assert np.array_equal(np.asarray(r["humans_3d"]).reshape(n, 15, 3), out["pose3d"][f, :n])


def test_bicubic_restatements_agree(oracle_lib):
    from oracle import decode_np
    rng = np.random.default_rng(3)
    for shape in ((5, 5), (3, 4), (28, 28)):
        src = rng.random(shape, dtype=np.float32)
        assert np.array_equal(decode_np.bicubic_upsample(src), oracle_lib.bicubic_upsample(src))


def test_bicubic_matches_opencv_cpp_path():
    """The third-party arithmetic itself: only where OpenCV is importable (it is in this image)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import decode_np
    rng = np.random.default_rng(5)
    prev = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        for shape in ((5, 5), (4, 5), (28, 28)):
            src = rng.random(shape, dtype=np.float32)
            ref = cv2.resize(src, None, fx=8, fy=8, interpolation=cv2.INTER_CUBIC)
            assert np.array_equal(decode_np.bicubic_upsample(src), ref)
    finally:
        cv2.ipp.setUseIPP(prev)


# ------------------------------------------------------------------------------------------------
# evaluator: host logic (popnet_b200.evaluate) + C oracle against the reference's numbers
# ------------------------------------------------------------------------------------------------
def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.fixture()
def eval_on_oracle(oracle_backend, monkeypatch):
    monkeypatch.setattr(E, "_backend", oracle_backend)
    return E


@pytest.mark.parametrize("tag,N,seed", [("small", 400, 7), ("c3", 4000, 0)])
def test_evaluator_matches_reference(tag, N, seed, eval_on_oracle):
    import warnings
    warnings.simplefilter("ignore")
    g = golden("eval_golden")
    ds = synth.eval_set(N, seed=seed)
    flat = [np.asarray([h for fr in ds[k] for h in fr], np.float64) for k in ("pred2d", "pred3d", "conf", "gt2d", "gt3d")]
    assert helpers.sha(*flat) == str(g[tag + "/sha"]), "synthetic eval set drifted from the fixture"
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
    assert np.array_equal(ap2, g[tag + "/ap2d"]) and np.array_equal(ap3, g[tag + "/ap3d"])
    # the integer contract
    m = E.match_counts(ds["pred2d"], ds["gt2d"], num_joints=15, dist_th=th2d)
    assert np.array_equal(m["hit_cnt"], g[tag + "/hit_pck2d"]) and np.array_equal(m["valid_cnt"], g[tag + "/valid2d"])
    assert m["samples_cnt"] == int(g[tag + "/samples"])
    m3 = E.match_counts(ds["pred2d"], ds["gt2d"], pred3d=ds["pred3d"], gt3d=ds["gt3d"], num_joints=15, dist_th=0.1)
    assert np.array_equal(m3["hit_cnt"], g[tag + "/hit_pck3d"]) and np.array_equal(m3["valid_cnt"], g[tag + "/valid3d"])
    for dim, c in ((2, c2), (3, c3)):
        assert np.array_equal(c["n_pos"], g[tag + "/map%d_npos" % dim])
        assert np.array_equal(c["n_gt"], g[tag + "/map%d_ngt" % dim])
        assert c["labels"].shape[0] == int(g[tag + "/map%d_nscores" % dim][0])
    if tag == "small":
        assert np.array_equal(m["dists"], g["small/dists2d"]) and np.array_equal(m3["dists"], g["small/dists3d"])
        assert np.array_equal(c2["labels"], g["small/map2_labels"]) and np.array_equal(c3["labels"], g["small/map3_labels"])


def test_evaluator_edge_cases(eval_on_oracle):
    """Edge cases the reference's code paths define: empty prediction list, frame without GT,
    missing joints, a prediction with no valid joint (whole frame unmatched), degenerate boxes (NaN IoU)."""
    K = 15
    base = np.stack([np.linspace(100, 200, K), np.linspace(50, 400, K)], 1)
    gt = [[base.tolist()], [], [base.tolist(), (base + 150).tolist()], [base.tolist()], [base.tolist()]]
    allmiss = (-np.ones((K, 2))).tolist()
    one = -np.ones((K, 2)); one[3] = [120.0, 80.0]
    gt_single = -np.ones((K, 2)); gt_single[3] = [120.0, 80.0]
    pred = [[], [], [(base + 3).tolist(), (base + 149).tolist()], [(base + 1).tolist(), allmiss], [one.tolist()]]
    gt[4] = [gt_single.tolist()]     # single-joint GT and pred: 0/0 IoU = NaN, which the reference treats as a match
    a, k = eval_on_oracle.eval_human_dataset_2d(pred, gt, K, 10.0, 0.5)
    m = eval_on_oracle.match_counts(pred, gt, num_joints=K, dist_th=10.0)
    assert m["samples_cnt"] == 5
    assert m["matched_pred"].tolist() == [-1, 0, 1, -1, 0]
    # frame 3: one prediction has no valid joint -> compute_bbox_from_humans returns [] -> nobody matches
    assert (m["dists"][3] == -1).all()
    # frame 4: NaN IoU matches; only joint 3 is valid in the prediction
    assert m["dists"][4][3] == 0.0 and (np.delete(m["dists"][4], 3) == -1).all()
    # a GT human with no valid joint makes the reference raise IndexError
    bad_gt = [[allmiss]]
    with pytest.raises(IndexError):
        eval_on_oracle.eval_human_dataset_2d([[base.tolist()]], bad_gt, K, 10.0, 0.5)
    # mAP: predictions in a frame without GT make the reference raise ValueError
    with pytest.raises(ValueError):
        _quiet(eval_on_oracle.eval_ap_3D, [[np.zeros((K, 3)).tolist()]], [], [[]], [], list(JOINT_NAMES), 0.1)


def test_packed_inputs_equal_list_inputs(oracle_lib, monkeypatch):
    """The evaluator accepts CSR-packed humans (evaluate.Packed) in place of the reference's ragged lists: same counters,
    same distances, same AP; the C list packer (csrc/packlists.c) reproduces np.asarray on the nested lists."""
    import contextlib
    import copy
    import io
    from oracle.backend import OracleBackend
    from popnet_b200 import evaluate as E
    from popnet_b200 import synth
    from popnet_b200.topology import JOINT_NAMES
    monkeypatch.setattr(E, "_backend", OracleBackend())
    ds = synth.eval_set(300, seed=12)
    ds["pred2d"][5] = []
    ds["pred3d"][5] = []
    ds["conf"][5] = []
    pk = {k: E.Packed(*E.pack_humans(ds[k], 15, 3 if "3d" in k else 2)) for k in ("pred2d", "pred3d", "gt2d", "gt3d")}
    flat = np.asarray([h for fr in ds["pred3d"] for h in fr], np.float64)
    assert np.array_equal(pk["pred3d"].flat, flat) and pk["pred3d"].off[-1] == len(flat)
    conf = E.Packed(E._pack_rows(ds["conf"], 15, np.float64), pk["pred2d"].off)
    names = list(JOINT_NAMES)
    with contextlib.redirect_stdout(io.StringIO()):
        a = E.eval_human_dataset_3d(ds["pred2d"], ds["gt2d"], ds["pred3d"], ds["gt3d"], 15, 0.1, 0.5)
        b = E.eval_human_dataset_3d(pk["pred2d"], pk["gt2d"], pk["pred3d"], pk["gt3d"], 15, 0.1, 0.5)
        c = E.eval_human_dataset_2d_PCKh(ds["pred2d"], ds["gt2d"], 0, 1, 15, 0.5, 0.5)
        d = E.eval_human_dataset_2d_PCKh(pk["pred2d"], pk["gt2d"], 0, 1, 15, 0.5, 0.5)
        e = E.eval_ap_3D(ds["pred3d"], copy.deepcopy(ds["conf"]), ds["gt3d"], [], names, 0.1)
        f = E.eval_ap_3D(pk["pred3d"], conf, pk["gt3d"], [], names, 0.1)
        g2 = E.eval_ap_mpii_v2(ds["pred2d"], copy.deepcopy(ds["conf"]), ds["gt2d"], [], 0, 1, names, 0.5)
        h2 = E.eval_ap_mpii_v2(pk["pred2d"], conf, pk["gt2d"], [], 0, 1, names, 0.5)
    for x, y in ((a, b), (c, d)):
        assert np.array_equal(np.asarray(x[0]), np.asarray(y[0])) and np.array_equal(np.asarray(x[1]), np.asarray(y[1]))
    assert np.array_equal(e, f) and np.array_equal(g2, h2)
    with pytest.raises(ValueError):
        E.pack_humans([[[[0.0, 1.0]] * 14]], 15, 2)
