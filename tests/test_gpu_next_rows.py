"""GPU tests of the "next" rows: device preprocessing, the stand-alone depth lift helper and the JSON wire
formats driving the evaluator."""
import json

import numpy as np
import pytest

import helpers

from popnet_b200 import synth
from popnet_b200.topology import ITOP, MP3DHP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cam,shape", [(MP3DHP, (512, 480)), (ITOP, (240, 320))], ids=["mp3dhp", "itop"])
def test_preprocess_matches_opencv(cam, shape, cuda_backend):
    cv2 = pytest.importorskip("cv2")
    from popnet_b200 import preprocess
    rng = np.random.default_rng(0)
    raw = rng.uniform(0.0, cam.depth_max + 1.0, (3,) + shape).astype(np.float32)
    raw[rng.random(raw.shape) < 0.04] = 0.0
    got = preprocess.preprocess_depth(raw, cam, 224).cpu().numpy()
    prev = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)          # OpenCV's own C++ bilinear path
        for b in range(3):
            img = cv2.resize(raw[b], (224, 224), interpolation=cv2.INTER_LINEAR)
            want = ((np.clip(img, 0, cam.depth_max) - cam.depth_mean) / cam.depth_std).astype(np.float32)
            assert np.abs(got[b, 0] - want).max() <= 2e-6
    finally:
        cv2.ipp.setUseIPP(prev)


def test_retrieve_depth_heat_weighted_matches_oracle(cuda_backend):
    from oracle import decode_np
    from popnet_b200.decode import retrieve_depth_heat_weighted
    rng = np.random.default_rng(1)
    heat = rng.random((28, 28), dtype=np.float32)
    depth = (rng.random((28, 28), dtype=np.float32) * 4 + 1).astype(np.float32)
    for c in [(0, 0), (27, 27), (0, 13), (5, 27), (12, 9)]:
        got = retrieve_depth_heat_weighted(c, depth, heat, radius=1)
        want = decode_np.retrieve_depth_heat_weighted(c, depth, heat, 1)
        assert got == want, (c, got, want)


def test_json_wire_formats_round_trip(tmp_path, cuda_backend):
    from popnet_b200 import io as pio
    ds = synth.eval_set(60, seed=21)
    labels = {"intrinsics": {"fx": MP3DHP.fx}}
    for i, (g2, g3) in enumerate(zip(ds["gt2d"], ds["gt3d"])):
        labels["%06d" % i] = [{"2d_joints": a, "3d_joints": b} for a, b in zip(g2, g3)]
    res = {"human_pred_set_2d": ds["pred2d"], "human_pred_set_3d": ds["pred3d"], "human_pred_set_part_conf": ds["conf"]}
    lp, rp = tmp_path / "labels.json", tmp_path / "results.json"
    lp.write_text(json.dumps(labels))
    rp.write_text(json.dumps(res))
    out = pio.evaluate_mp_human_3d(str(lp), str(rp))
    # same numbers as calling the evaluator directly on the lists
    from popnet_b200 import evaluate as E
    _, k2 = E.eval_human_dataset_2d_PCKh(ds["pred2d"], ds["gt2d"], 0, 1, 15, 0.5, 0.5)
    assert np.array_equal(np.asarray(out["pckh_2d"]), np.asarray(k2))
    assert 0.0 < out["overall"]["pckh_2d"] <= 1.0 and 0.0 < out["overall"]["map_3d"] <= 100.0


def test_depth_read_variants_match_oracle(cuda_backend):
    """retrieve_depth_weighted / retrieve_depth_heat_max (lib/utils/common.py:251-318), SURVEY.md 8(f) row 4."""
    from oracle import decode_np
    from popnet_b200.decode import retrieve_depth_weighted, retrieve_depth_heat_max
    rng = np.random.default_rng(3)
    heat = (rng.random((28, 28), dtype=np.float32) - 0.2).astype(np.float32)        # some negative cells: clamp path
    heat[10:13, 4:7] = 0.5                                                            # a plateau: first maximum wins
    depth = (rng.random((28, 28), dtype=np.float32) * 4 + 1).astype(np.float32)
    for c in [(0, 0), (27, 27), (0, 13), (5, 27), (12, 9), (5, 11), (26, 1)]:
        assert retrieve_depth_weighted(c, depth, radius=1) == decode_np.retrieve_depth_weighted(c, depth, 1), c
        assert retrieve_depth_heat_max(c, depth, heat.copy(), radius=1) == decode_np.retrieve_depth_heat_max(c, depth, heat, 1), c
    # batched, all cells of a plane, against the restatement
    q = np.array([[0, x, y] for y in range(28) for x in range(28)], np.int32)
    for mode, fn in ((1, lambda c: decode_np.retrieve_depth_weighted(c, depth, 1)),
                     (2, lambda c: decode_np.retrieve_depth_heat_max(c, depth, heat, 1))):
        got = cuda_backend.lift_depth(heat[None], depth[None], q, mode=mode)
        want = np.array([fn((int(x), int(y))) for _, x, y in q], np.float32)
        assert np.array_equal(got, want), mode


@pytest.mark.parametrize("radius", [0, 2, 3, 5])
def test_depth_reads_other_radii_match_oracle(radius, cuda_backend):
    """The helpers' `radius` argument (common.py:251,272,296): windows of 1 .. 121 cells, every cell of the grid (clipped
    windows at the borders), all three reads, bit-exact against the restatement (which test_refcheck pins to the reference)."""
    from oracle import decode_np
    from popnet_b200.decode import retrieve_depth_heat_weighted
    rng = np.random.default_rng(40 + radius)
    heat = (rng.random((28, 28), dtype=np.float32) - 0.2).astype(np.float32)
    heat[10:13, 4:7] = 0.5
    depth = (rng.random((28, 28), dtype=np.float32) * 4 + 1).astype(np.float32)
    q = np.array([[0, x, y] for y in range(28) for x in range(28)], np.int32)
    for mode, fn in ((0, lambda c: decode_np.retrieve_depth_heat_weighted(c, depth, heat, radius)),
                     (1, lambda c: decode_np.retrieve_depth_weighted(c, depth, radius)),
                     (2, lambda c: decode_np.retrieve_depth_heat_max(c, depth, heat, radius))):
        got = cuda_backend.lift_depth(heat[None], depth[None], q, mode=mode, radius=radius)
        want = np.array([fn((int(x), int(y))) for _, x, y in q], np.float32)
        assert np.array_equal(got, want), (mode, radius)
    assert retrieve_depth_heat_weighted((3, 4), depth, heat.copy(), radius=radius) == \
        decode_np.retrieve_depth_heat_weighted((3, 4), depth, heat, radius)
    with pytest.raises(ValueError):
        retrieve_depth_heat_weighted((3, 4), depth, heat, radius=6)


def test_decode_coco_topology_vs_oracle(cuda_backend, oracle_lib):
    """The decode kernels take the skeleton as data: COCO's 18 keypoints / 19 limbs (pafprocess.h:21-24), byte-for-byte
    against the C oracle (which tests/test_refcheck.py pins to the reference with the same topology)."""
    from popnet_b200 import _abi, synth, topology
    heat, paf, depth, _ = synth.map_batch(48, seed=31, persons=(1, 7), noise=0.01, limbs=topology.COCO_LIMBS,
                                          template=synth._TEMPLATE_COCO)
    params = _abi.make_decode_params(topology.coco_config(), topology.MP3DHP)
    dev = cuda_backend.decode(heat, paf, depth, params)
    ora = oracle_lib.decode(heat, paf, depth, params)
    assert helpers.records_equal(dev, ora) == []
    assert int(dev["n_person"].sum()) >= 48 and not dev["flags"].any()
