"""Forward parity.

CPU part: the fp32 torch oracle (oracle/forward_torch.py) against the reference module's own outputs
(golden fixture), and the state-dict contract of the host mirror.
GPU part: the CUDA forward (tcgen05 path and the CUDA-core check path, both through popnet_forward)
against the oracle; tolerance max-abs <= 1e-2 on the three output maps (BASELINE.json north_star).
"""
import numpy as np
import pytest
import torch

import helpers
from helpers import golden
from popnet_b200 import network

TOL = 1e-2


def _state(style):
    sd = network.synth_state_dict(seed=11, style=style)
    g = golden("forward_golden")
    assert helpers.sha(*[sd[k] for k in sorted(sd)]) == str(g[style + "/state_sha"]), "synthetic checkpoint drifted"
    return sd


@pytest.mark.parametrize("style", ["reference", "trained_like"])
def test_torch_oracle_matches_reference_module(style):
    from oracle import forward_torch
    g = golden("forward_golden")
    x = g["x"].astype(np.float32)
    (paf, heat, depth), saved = forward_torch.forward(_state(style), x)
    for name, t in (("paf", paf), ("heat", heat), ("depth", depth), ("paf1", saved[0]), ("heat1", saved[1]), ("depth1", saved[2])):
        err = np.abs(t.numpy() - g["%s/%s" % (style, name)]).max()
        assert err < 2e-5, (name, err)      # same fp32 ops; thread-count dependent summation order only


def test_state_dict_contract():
    """234 reference keys, DataParallel prefix accepted, canonical conv order has 39 layers."""
    m = network.rtpose_light3d(15, 14, 2, input_dim=1)
    sd = m.state_dict()
    assert len(sd) == 234
    assert "model0.layer2.0.downsample.1.running_var" in sd and "model2_3.12.bias" in sd and "model1_1.1.num_batches_tracked" in sd
    m.load_state_dict({"module." + k: v for k, v in sd.items()})
    assert len(m.conv_layers()) == 39
    assert sum(p.numel() for p in m.parameters()) == 5525814
    with pytest.raises(ValueError):
        network.rtpose_light3d(15, 14, 3, input_dim=1)


def test_forward_requires_cuda_tensor():
    from popnet_b200._lib import PopnetError
    m = network.rtpose_light3d(15, 14, 2, input_dim=1)
    with pytest.raises(PopnetError):
        m(torch.zeros(1, 1, 224, 224))


# Operand format vs tolerance (measured): fp16 operands -- the product default -- hold the north
# star's 1e-2 bound on every checkpoint (reference init, the He-scaled stress fixture, the trained fixture checkpoint);
# bf16 operands (selectable, same speed) hold it on reference-initialised weights only: their 8-bit mantissa gives
# 2.6e-2 .. 3.6e-2 on trained weights.  The bound is asserted for every combination that is claimed; the bf16 rows on
# O(1)-activation weights assert only that the result is sane (finite, within 6e-2) and print the measured error.
TOLS = {("bf16", "reference"): 1e-2, ("fp16", "reference"): 1e-2, ("fp16", "trained_like"): 1e-2}
BF16_SANITY = 6e-2


@pytest.mark.gpu
@pytest.mark.parametrize("style", ["reference", "trained_like"])
@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_cuda_forward_vs_oracle(style, impl, dtype, cuda_backend):
    from oracle import forward_torch
    from popnet_b200 import _abi
    g = golden("forward_golden")
    x = g["x"].astype(np.float32)
    sd = _state(style)
    m = network.rtpose_light3d(15, 14, 2, input_dim=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.impl = _abi.FWD_IMPL_SIMT if impl == "simt" else _abi.FWD_IMPL_TCGEN05
    m.operand_dtype = _abi.OPERAND_BF16 if dtype == "bf16" else _abi.OPERAND_FP16
    tol = TOLS.get((dtype, style), BF16_SANITY)
    (paf, heat, depth), saved = m(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    (opaf, oheat, odepth), osaved = forward_torch.forward(sd, x)
    errs = {}
    for name, a, b in (("paf", paf, opaf), ("heat", heat, oheat), ("depth", depth, odepth),
                       ("paf1", saved[0], osaved[0]), ("heat1", saved[1], osaved[1]), ("depth1", saved[2], osaved[2])):
        errs[name] = float((a.cpu() - b).abs().max())
        # and against the reference module's own numbers
        errs[name + "_vs_ref"] = float(np.abs(a.cpu().numpy() - g["%s/%s" % (style, name)]).max())
    print("max-abs errors (%s, %s, %s):" % (style, impl, dtype), {k: round(v, 5) for k, v in errs.items()})
    assert all(np.isfinite(v) and v <= tol for v in errs.values()), errs


@pytest.mark.gpu
def test_cuda_forward_batch_invariance(cuda_backend):
    """Frames are independent: a frame's maps do not depend on its position in the batch or on batch size."""
    from popnet_b200 import synth
    sd = network.synth_state_dict(seed=11, style="trained_like")
    m = network.rtpose_light3d(15, 14, 2, input_dim=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    x = torch.from_numpy(synth.depth_frames(9, seed=99)).cuda()
    (p9, h9, d9), _ = m(x)
    (p1, h1, d1), _ = m(x[4:5])
    (p3, h3, d3), _ = m(x[3:6])
    torch.cuda.synchronize()
    assert torch.equal(p9[4:5], p1) and torch.equal(h9[4:5], h1) and torch.equal(d9[4:5], d1)
    assert torch.equal(p9[3:6], p3) and torch.equal(h9[3:6], h3) and torch.equal(d9[3:6], d3)


def _maps_for_tunings(tunings, batches=((3, 5), (64, 6))):
    """The six output maps at every (batch, seed) for every schedule in `tunings` (name -> PopnetNetConfig.tuning bits)."""
    from popnet_b200 import synth
    sd = network.synth_state_dict(seed=11, style="trained_like")
    outs = {}
    for tag, bits in tunings.items():
        m = network.rtpose_light3d(15, 14, 2, input_dim=1)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        m.tuning = bits
        res = []
        for B, seed in batches:
            x = torch.from_numpy(synth.depth_frames(B, seed=seed)).cuda()
            (p, h, d), saved = m(x)
            torch.cuda.synchronize()
            res += [p.clone(), h.clone(), d.clone()] + [t.clone() for t in saved[:3]]
        outs[tag] = res
    return outs


@pytest.mark.gpu
def test_cuda_forward_layer_chain_identical(cuda_backend):
    """POPNET_TUNE_CHAIN runs the four 64 -> 64 layers of the 112 x 112 block as ONE spatially pipelined launch (CTA
    slices linked by per-tile progress counters, conv_kernels.cu "CHAIN").  Same kernel body and K order per output as the
    default one-launch-per-layer schedule (with and without the zig-zag order): bit-identical maps, at a batch with few
    tiles per CTA slice (3), at the bench batch (64), and repeated (the counters are re-zeroed per forward)."""
    from popnet_b200 import _abi
    outs = _maps_for_tunings({"chain": _abi.TUNE_CHAIN, "chain-again": _abi.TUNE_CHAIN, "layers": 0,
                              "layers-nozz": _abi.TUNE_NO_ZIGZAG, "chain-nozz": _abi.TUNE_CHAIN | _abi.TUNE_NO_ZIGZAG},
                             batches=((3, 5), (64, 6), (1, 7), (64, 8)))
    for tag in ("chain-again", "layers", "layers-nozz", "chain-nozz"):
        for a, b in zip(outs["chain"], outs[tag]):
            assert torch.equal(a, b), tag


@pytest.mark.gpu
def test_cuda_forward_multicast_variant_identical(cuda_backend):
    """POPNET_TUNE_MC runs the N = 256 stage layers as cluster-of-two kernels whose weight stages are multicast (128-position
    tiles, dummy tile slots, cross-CTA stage release).  Same K order per output, so the maps must be bit-identical to the
    default path -- at a batch with dummy slots (odd tile counts) and at the bench batch."""
    from popnet_b200 import _abi
    outs = _maps_for_tunings({"0": 0, "1": _abi.TUNE_MC})
    for a, b in zip(outs["0"], outs["1"]):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_cuda_forward_pair_kernel_identical(cuda_backend):
    """The cta_group::2 pair kernel (opt-in, POPNET_TUNE_PAIR(4|3)) for the 64 -> 64 layers at 112 x 112 against the single-CTA
    kernel, without and with the residual layers, 512- and 384-position tiles: same K order per output -> bit-identical.
    Batch 3 gives odd tile counts (dummy tile slots), batch 64 is the bench shape."""
    from popnet_b200 import _abi
    nc = 0
    outs = _maps_for_tunings({"off": nc, "default": nc | _abi.tune_pair(4), "res3": nc | _abi.tune_pair(3) | _abi.TUNE_PAIR_RES,
                              "res4": nc | _abi.tune_pair(4) | _abi.TUNE_PAIR_RES})
    for tag in ("default", "res3", "res4"):
        for a, b in zip(outs["off"], outs[tag]):
            assert torch.equal(a, b), tag


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(160, 192), (224, 96), (64, 64)], ids=lambda v: "%dx%d" % v)
def test_cuda_forward_other_input_sizes(hw, cuda_backend):
    """The layer plan is generic in the input size (multiples of 8, non-square included): rows / columns / image pitch
    all differ here, so any place that confuses H with W or Hs with Wp fails against the fp32 oracle."""
    from oracle import forward_torch
    from popnet_b200 import _abi
    sd = network.synth_state_dict(seed=11, style="trained_like")
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    x = rng.normal(0.0, 1.0, (3, 1, hw[0], hw[1])).astype(np.float32)
    m = network.rtpose_light3d(15, 14, 2, input_dim=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.operand_dtype = _abi.OPERAND_FP16
    (paf, heat, depth), saved = m(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    (opaf, oheat, odepth), osaved = forward_torch.forward(sd, x)
    assert tuple(paf.shape) == (3, 28, hw[0] // 8, hw[1] // 8)
    for name, a, b in (("paf", paf, opaf), ("heat", heat, oheat), ("depth", depth, odepth), ("paf1", saved[0], osaved[0])):
        err = float((a.cpu() - b).abs().max())
        assert np.isfinite(err) and err <= TOL, (name, err)


@pytest.mark.gpu
def test_cuda_forward_schedule_switches_identical(cuda_backend):
    """The round-2 launch-schedule switches -- SM reserve, balanced grids, 256-position stage tiles, prologue prefill off,
    cluster-of-two stage launches (with and without the multicast kernels) -- only change WHO computes a tile and when:
    the maps must be bit-identical to the default schedule, at a batch with few tiles (3: odd tile counts, CTAs without
    tiles in the cluster launches) and at the bench batch."""
    from popnet_b200 import _abi
    tun = {"default": 0, "reserve8": _abi.TUNE_RESERVE_SMS(2), "reserve28": _abi.TUNE_RESERVE_SMS(7),
           "balance": _abi.TUNE_BALANCE, "nacc2-balance": _abi.TUNE_BALANCE | _abi.tune_stage_nacc(2),
           "no-prefill": _abi.TUNE_NO_PREFILL, "cluster-all": _abi.TUNE_CLUSTER_ALL,
           "cluster-all-mc-reserve": _abi.TUNE_CLUSTER_ALL | _abi.TUNE_MC | _abi.TUNE_RESERVE_SMS(2)}
    outs = _maps_for_tunings(tun)
    for tag in tun:
        for a, b in zip(outs["default"], outs[tag]):
            assert torch.equal(a, b), tag
