"""GPU parity: CUDA decode + lift (through the C ABI) against the C oracle and the reference's golden outputs."""
import numpy as np
import pytest

import helpers
from helpers import DECODE_CASES, golden
from popnet_b200 import synth

pytestmark = pytest.mark.gpu

# popnet_decode computes every record byte for byte like the oracle whatever the CTA limit of its three persistent kernels:
# 0 = all SMs (a stand-alone call), 8 = the pipelined step's setting (the SMs the convolution grids leave free), 1 = a single
# CTA walks every item (the longest per-CTA item lists, the big-item pass of the limb kernel over many items).
SCHEDULES = [("all_sms", 0), ("8_ctas", 8), ("1_cta", 1)]
schedules = pytest.mark.parametrize("schedule", [v for _, v in SCHEDULES], ids=[n for n, _ in SCHEDULES])


@schedules
@pytest.mark.parametrize("case", DECODE_CASES, ids=[c[0] for c in DECODE_CASES])
def test_decode_bitwise_vs_oracle_and_reference(case, schedule, cuda_backend, oracle_lib):
    g = golden("decode_golden")
    heat, paf, depth = helpers.decode_case_inputs(case)
    params = helpers.params_for(case[5], max_ctas=schedule)
    dev = cuda_backend.decode(heat, paf, depth, params)
    ora = oracle_lib.decode(heat, paf, depth, params)
    assert helpers.records_equal(dev, ora) == []
    for f in range(heat.shape[0]):
        ok, why = helpers.compare_to_golden(dev, f, g, "%s/native/%d/" % (case[0], f), exact=True)
        assert ok, "frame %d vs reference (OpenCV C++ path): %s" % (f, why)
        ok, why = helpers.compare_to_golden(dev, f, g, "%s/ipp/%d/" % (case[0], f), exact=False)
        assert ok, "frame %d vs reference (IPP path): %s" % (f, why)


@pytest.mark.parametrize("batch,persons,seed", [(64, (1, 6), 1234), (256, (12, 16), 4242), (512, (1, 6), 77)],
                         ids=["C2-b64", "C5-crowd-b256", "C4-b512"])
@schedules
def test_decode_full_size_vs_oracle(batch, persons, seed, schedule, cuda_backend, oracle_lib):
    """BASELINE.json configs C2 / C5 / C4 at their full batch sizes, byte-for-byte against the oracle."""
    heat, paf, depth, _ = synth.map_batch(batch, seed=seed, persons=persons, noise=0.01)
    params = helpers.params_for("MP3DHP", max_ctas=schedule)
    dev = cuda_backend.decode(heat, paf, depth, params)
    ora = oracle_lib.decode(heat, paf, depth, params)
    assert helpers.records_equal(dev, ora) == []
    assert int(dev["n_person"].sum()) > batch * persons[0] * 0.5


def test_decode_shard_invariance(cuda_backend):
    """Frames are independent: decoding a batch in shards gives the same records (the multi-GPU contract)."""
    heat, paf, depth, _ = synth.map_batch(48, seed=9, persons=(1, 6), noise=0.01)
    params = helpers.params_for("MP3DHP")
    full = cuda_backend.decode(heat, paf, depth, params)
    parts = [cuda_backend.decode(heat[s], paf[s], depth[s], params) for s in (slice(0, 16), slice(16, 17), slice(17, 48))]
    cat = {k: np.concatenate([p[k] for p in parts], 0) for k in full}
    assert helpers.records_equal(full, cat) == []


@schedules
def test_decode_degenerate_overflow_flags(schedule, cuda_backend, oracle_lib):
    """Untrained-network-like maps (heat ~ 0.5 everywhere): hundreds of plateau peaks per joint
    type.  Capacities overflow; flags and the truncated records must match the oracle's."""
    rng = np.random.default_rng(0)
    heat = (0.5 + 0.01 * rng.standard_normal((3, 16, 28, 28))).astype(np.float32)
    heat[2] = 0.5                                         # one giant plateau: every cell is a peak
    paf = (0.05 * rng.standard_normal((3, 28, 28, 28))).astype(np.float32)
    depth = rng.standard_normal((3, 15, 28, 28)).astype(np.float32)
    params = helpers.params_for("MP3DHP", max_ctas=schedule)
    dev = cuda_backend.decode(heat, paf, depth, params)
    ora = oracle_lib.decode(heat, paf, depth, params)
    assert (dev["flags"] & 1).all()
    assert helpers.records_equal(dev, ora) == []


def test_paf_to_pose_signature(cuda_backend):
    """The reference-facing call: HWC maps of one frame in, (joint_list, person_to_joint_assoc) out."""
    from types import SimpleNamespace as NS
    from popnet_b200.decode import paf_to_pose, paf_to_human_list
    g = golden("decode_golden")
    heat, paf, depth = helpers.decode_case_inputs(DECODE_CASES[0])
    cfg = NS(MODEL=NS(NUM_KEYPOINTS=15, NUM_LIMBS=14, DOWNSAMPLE=8),
             TEST=NS(THRESH_HEATMAP=0.1, THRESH_PAF=0.05, NUM_INTERMED_PTS_BETWEEN_KEYPOINTS=10))
    for f in (0, 1, 5):
        jl, assoc = paf_to_pose(heat[f].transpose(1, 2, 0), paf[f].transpose(1, 2, 0), cfg)
        assert np.array_equal(jl, g["mp/native/%d/joint_list" % f])
        assert np.array_equal(np.asarray(assoc).reshape(-1, 17), g["mp/native/%d/assoc" % f])
        humans, vis, conf = paf_to_human_list(jl, assoc)
        assert len(humans) == len(g["mp/native/%d/assoc" % f])


@pytest.mark.parametrize("size", [160, 320], ids=["grid20", "grid40"])
def test_decode_other_grid_sizes_vs_oracle(size, cuda_backend, oracle_lib):
    """The decode takes the grid size as data (up to 64 x 64 cells): 20 x 20 and 40 x 40 grids, byte-for-byte vs the oracle."""
    from popnet_b200 import _abi
    from popnet_b200.topology import DecodeConfig, MP3DHP
    heat, paf, depth, _ = synth.map_batch(24, seed=5 + size, persons=(1, 5), noise=0.01, size=size)
    assert heat.shape[-1] == size // 8
    params = _abi.make_decode_params(DecodeConfig(), MP3DHP, input_size=size)
    dev = cuda_backend.decode(heat, paf, depth, params)
    ora = oracle_lib.decode(heat, paf, depth, params)
    assert helpers.records_equal(dev, ora) == []
    assert int(dev["n_person"].sum()) >= 24 * 0.5


@pytest.mark.parametrize("rows,cols", [(20, 28), (28, 17), (9, 40)], ids=lambda v: str(v))
def test_decode_non_square_grid_vs_oracle(rows, cols, cuda_backend, oracle_lib):
    """Rows != columns (the reference's paf_to_pose takes any H x W, paf_to_pose.py:354-377): crops of rendered maps,
    byte-for-byte vs the C oracle (pinned to the reference on the same crops by tests/test_refcheck.py), through both the
    batched entry and the reference-signature paf_to_pose (whose grid comes from the maps, not from a square size)."""
    from types import SimpleNamespace as NS
    from popnet_b200 import _abi
    from popnet_b200.decode import paf_to_pose, records_to_reference
    from popnet_b200.topology import DecodeConfig, MP3DHP
    heat, paf, depth, _ = synth.map_batch(12, seed=600 + rows, persons=(2, 6), noise=0.01, size=384)
    heat, paf, depth = (np.ascontiguousarray(t[:, :, 3:3 + rows, 5:5 + cols]) for t in (heat, paf, depth))
    params = _abi.make_decode_params(DecodeConfig(), MP3DHP, input_size=224, grid_hw=(rows, cols))
    dev = cuda_backend.decode(heat, paf, depth, params)
    ora = oracle_lib.decode(heat, paf, depth, params)
    assert helpers.records_equal(dev, ora) == []
    assert int(dev["peak_count"].sum()) > 12 * 10
    cfg = NS(MODEL=NS(NUM_KEYPOINTS=15, NUM_LIMBS=14, DOWNSAMPLE=8),
             TEST=NS(THRESH_HEATMAP=0.1, THRESH_PAF=0.05, NUM_INTERMED_PTS_BETWEEN_KEYPOINTS=10))
    for f in (0, 7):
        jl, assoc = paf_to_pose(heat[f].transpose(1, 2, 0), paf[f].transpose(1, 2, 0), cfg)
        ojl, oassoc = records_to_reference(ora, f, 15)
        assert np.array_equal(jl, ojl) and np.array_equal(np.asarray(assoc), np.asarray(oassoc))


def test_decode_rejects_mismatched_shapes(cuda_backend):
    """The C ABI takes raw pointers; the Python boundary refuses maps whose shape disagrees with the parameter block
    instead of indexing them with the wrong pitch."""
    from popnet_b200 import _abi
    from popnet_b200._lib import PopnetError
    from popnet_b200.topology import DecodeConfig, MP3DHP
    heat, paf, depth, _ = synth.map_batch(2, seed=1, persons=(1, 2))
    params = _abi.make_decode_params(DecodeConfig(), MP3DHP)
    with pytest.raises(PopnetError):
        cuda_backend.decode(heat[:, :, :20], paf[:, :, :20], depth[:, :, :20], params)          # 20 x 28 maps, 28 x 28 params
    with pytest.raises(PopnetError):
        cuda_backend.decode(heat, paf, np.concatenate([depth, depth[:, :5]], 1), params)          # 20 depth planes, K = 15
    p20 = _abi.make_decode_params(DecodeConfig(), MP3DHP, depth_channels=20)
    d20 = np.ascontiguousarray(np.concatenate([depth, depth[:, :5]], 1))
    a, b = cuda_backend.decode(heat, paf, d20, p20), cuda_backend.decode(heat, paf, depth, params)
    assert helpers.records_equal(a, b) == []                                                     # joint j reads plane j


@schedules
def test_decode_odd_capacities(schedule, cuda_backend, oracle_lib):
    """Capacities that are not multiples of 4 / 8 (max_peaks 37, max_persons 21): every per-warp shared-memory slice of the
    decode kernels must stay aligned, and truncation / overflow flags must match the oracle's."""
    heat, paf, depth, _ = synth.map_batch(24, seed=909, persons=(8, 14), noise=0.02)
    params = helpers.params_for("MP3DHP", max_ctas=schedule, max_peaks=37, max_persons=21)
    dev = cuda_backend.decode(heat, paf, depth, params)
    ora = oracle_lib.decode(heat, paf, depth, params)
    assert helpers.records_equal(dev, ora) == []
