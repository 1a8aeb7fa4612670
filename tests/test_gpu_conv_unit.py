"""GPU unit test of the convolution kernels on random C8P tensors: tcgen05 kernel vs the CUDA-core check
kernel vs torch.nn.functional.conv2d (fp32 on the same bf16-rounded operands), for every instantiated
tile shape (NT, NACC, TAPS)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD, ROUND = 128, 512


class DebugConv(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("in_plane_stride", C.c_longlong), ("w", C.c_void_p), ("shift", C.c_void_p),
                ("out", C.c_void_p), ("out_plane_stride", C.c_longlong), ("res", C.c_void_p),
                ("res_plane_stride", C.c_longlong), ("head_out", C.c_void_p)] + \
               [(n, C.c_int) for n in ("P", "Hs", "Wp", "chunks", "a_stages", "act", "cout", "cout_pad", "nt", "nacc",
                                       "taps", "impl", "fmt", "dbg")] + [("probe", C.c_void_p)]


DT = torch.bfloat16


def to_c8p(x):
    """[N,C,H,W] fp32 -> C8P planes [C/8, plane_len, 8]: rows of W pixels + one zero cell, two zero rows on top and one
    zero row after every image (guards filled with NaN on purpose)."""
    N, Cc, H, W = x.shape
    P = (2 + N * (H + 1)) * (W + 1)
    plen = GUARD + (P + ROUND - 1) // ROUND * ROUND + 512 + GUARD
    xp = torch.zeros((Cc, 2 + N * (H + 1), W + 1), device=x.device)
    xp[:, 2:].view(Cc, N, H + 1, W + 1)[:, :, :H, :W] = x.permute(1, 0, 2, 3)
    planes = torch.full((Cc // 8, plen, 8), float("nan"), device=x.device, dtype=DT)
    v = xp.reshape(Cc // 8, 8, P).permute(0, 2, 1)
    planes[:, GUARD:GUARD + P] = v.to(DT)
    return planes, P, plen


def from_c8p(planes, N, Cc, H, W):
    """Returns (pixels [N,C,H,W], zero_cells): the second tensor gathers everything that must be zero."""
    P = (2 + N * (H + 1)) * (W + 1)
    v = planes[:, GUARD:GUARD + P].float().permute(0, 2, 1).reshape(Cc, 2 + N * (H + 1), W + 1)
    body = v[:, 2:].reshape(Cc, N, H + 1, W + 1)
    pix = body[:, :, :H, :W].permute(1, 0, 2, 3)
    zeros = torch.cat([v[:, :2].reshape(Cc, -1), body[:, :, H].reshape(Cc, -1), body[:, :, :H, W].reshape(Cc, -1)], 1)
    return pix, zeros


def pack_w(w, nt, cin_pad, cout_pad):
    """[cout,cin,k,k] fp32 -> [n_tile][tap][cin_pad/8][nt][8] bf16."""
    cout, cin, k, _ = w.shape
    wp = torch.zeros((cout_pad, cin_pad, k * k), device=w.device)
    wp[:cout, :cin] = w.reshape(cout, cin, k * k)
    v = wp.reshape(cout_pad // nt, nt, cin_pad // 8, 8, k * k).permute(0, 4, 2, 1, 3).contiguous()
    return v.to(DT)


CASES = [  # nt, nacc, taps, cin, cout, H, W, N, act, residual, head
    (64, 2, 9, 64, 64, 20, 24, 3, 1, True, False),
    (64, 4, 9, 128, 64, 12, 12, 5, 2, False, False),
    (64, 3, 9, 64, 64, 20, 24, 7, 1, True, False),
    (64, 4, 9, 64, 64, 12, 12, 9, 2, False, False),
    (128, 2, 9, 64, 128, 14, 10, 4, 1, False, False),
    (128, 1, 9, 64, 128, 14, 10, 9, 1, False, False),          # weights resident (layer2.0.conv1)
    (128, 4, 9, 192, 128, 12, 12, 5, 2, False, False),
    (128, 4, 1, 256, 128, 12, 12, 5, 2, False, False),
    (128, 4, 1, 64, 128, 14, 10, 4, 0, False, False),
    (256, 2, 9, 128, 256, 12, 12, 3, 2, False, False),
    (256, 2, 9, 256, 256, 28, 28, 2, 2, False, False),
    (32, 4, 1, 128, 28, 12, 12, 5, 3, False, True),
    (16, 4, 9, 128, 16, 12, 12, 5, 4, False, True),
    (16, 4, 9, 64, 15, 12, 12, 5, 3, False, True),
]


@pytest.mark.parametrize("fmt", [0, 1], ids=["bf16", "fp16"])
@pytest.mark.parametrize("case", CASES, ids=["nt%d-acc%d-t%d-cin%d-cout%d" % c[:5] for c in CASES])
def test_conv_tc_vs_simt_vs_torch(case, fmt, cuda_backend):
    nt, nacc, taps, cin, cout, H, W, N, act, use_res, head = case
    global DT
    DT = torch.bfloat16 if fmt == 0 else torch.float16
    lib = cuda_backend.lib
    lib.popnet_debug_conv.restype = C.c_int
    lib.popnet_debug_conv.argtypes = [C.POINTER(DebugConv), C.c_void_p]
    g = torch.Generator(device="cuda").manual_seed(1000 + cin + cout)
    k = 3 if taps == 9 else 1
    cin_pad, cout_pad = (cin + 63) // 64 * 64, (cout + nt - 1) // nt * nt
    x = torch.randn((N, cin_pad, H, W), device="cuda", generator=g)
    w = torch.randn((cout, cin, k, k), device="cuda", generator=g) * (1.0 / (cin * k * k)) ** 0.5
    shift = torch.randn((cout_pad,), device="cuda", generator=g) * 0.1
    shift[cout:] = 0
    res = torch.randn((N, cout_pad, H, W), device="cuda", generator=g) if use_res else None
    xin, P, plen = to_c8p(x)
    wpk = pack_w(w, nt, cin_pad, cout_pad)
    rin = to_c8p(res)[0] if use_res else None
    # fp32 reference on the bf16-rounded operands
    xr = x[:, :cin].to(DT).float()
    wr = w.to(DT).float()
    ref = torch.nn.functional.conv2d(xr, wr, None, 1, k // 2) + shift[:cout].view(1, -1, 1, 1)
    if use_res:
        ref = ref + res[:, :cout].to(DT).float()
    ref = {0: ref, 1: ref.relu(), 2: torch.nn.functional.leaky_relu(ref, 0.1), 3: (ref.sigmoid() - 0.5) * 4,
           4: ref.sigmoid()}[act]
    results = {}
    for name, impl in (("simt", 1), ("tc", 0)):
        out = torch.full((cout_pad // 8, plen, 8), float("nan"), device="cuda", dtype=DT)
        hout = torch.full((N, cout, H, W), float("nan"), device="cuda") if head else None
        d = DebugConv(inp=xin[:, GUARD:].data_ptr(), in_plane_stride=plen * 8, w=wpk.data_ptr(), shift=shift.data_ptr(),
                      out=out[:, GUARD:].data_ptr(), out_plane_stride=plen * 8,
                      res=(rin[:, GUARD:].data_ptr() if use_res else None), res_plane_stride=plen * 8,
                      head_out=(hout.data_ptr() if head else None), P=P, Hs=H + 1, Wp=W + 1, chunks=cin_pad // 64,
                      a_stages=2, act=act, cout=cout, cout_pad=cout_pad, nt=nt, nacc=nacc,
                      taps=taps, impl=impl, fmt=fmt)
        rc = lib.popnet_debug_conv(C.byref(d), None)
        assert rc == 0, rc
        torch.cuda.synchronize()
        o, zero_cells = from_c8p(out, N, cout_pad, H, W)
        assert torch.equal(zero_cells, torch.zeros_like(zero_cells)), "%s: zero cells of the layout not zero" % name
        assert torch.equal(o[:, cout:], torch.zeros_like(o[:, cout:])), "%s: padded channels not zero" % name
        results[name] = (o[:, :cout], hout)
        err = (results[name][0] - ref).abs().max().item()
        tol = 0.03 * max(1.0, ref.abs().max().item())       # bf16 output rounding
        assert err < tol, "%s vs torch: max-abs %.4g (tol %.3g)" % (name, err, tol)
        if head:
            herr = (hout - ref).abs().max().item()
            assert herr < 2e-3, "%s head fp32 vs torch: %.4g" % (name, herr)
    # same operands, same math up to summation order: the two kernels agree to bf16 resolution
    diff = (results["tc"][0] - results["simt"][0]).abs().max().item()
    assert diff <= 0.02 * max(1.0, ref.abs().max().item()), diff
