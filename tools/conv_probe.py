"""Time single convolution layers of the forward at full batch through the debug hook and print the
in-kernel phase stamps (clock64) of a few CTAs.  Tuning aid; not part of the product."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from popnet_b200 import _lib  # noqa: E402
from test_gpu_conv_unit import DebugConv, GUARD, ROUND  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--layers", default="128->128@28,256->256@28", help="comma-separated names from LAYERS_ALL")
ap.add_argument("--ctas", default="0", help="comma-separated CTA indices whose probes are printed")
a = ap.parse_args()
lib = _lib.get()
lib.popnet_debug_conv.restype = C.c_int
lib.popnet_debug_conv.argtypes = [C.POINTER(DebugConv), C.c_void_p]

# name, nt, nacc, taps, cin, cout, H
LAYERS_ALL = [("64->64@112", 64, 2, 9, 64, 64, 112), ("64->64@112 acc3", 64, 3, 9, 64, 64, 112), ("64->64@112 acc4", 64, 4, 9, 64, 64, 112),
          ("64->128@56", 128, 2, 9, 64, 128, 56), ("64->128@56 acc4", 128, 4, 9, 64, 128, 56),
          ("128->128@56 acc2", 128, 2, 9, 128, 128, 56), ("128->128@28 acc2", 128, 2, 9, 128, 128, 28),
          ("128->128@56", 128, 4, 9, 128, 128, 56), ("1x1 128->128@56", 128, 4, 1, 128, 128, 56),
          ("256->256@28", 256, 2, 9, 256, 256, 28), ("128->128@28", 128, 4, 9, 128, 128, 28),
          ("64->64@28", 64, 4, 9, 64, 64, 28)]
LAYERS = [l for l in LAYERS_ALL if l[0] in a.layers.split(",")]
for name, nt, nacc, taps, cin, cout, H in LAYERS:
    N = a.batch
    P = (2 + N * (H + 1)) * (H + 1)
    plen = GUARD + (P + ROUND - 1) // ROUND * ROUND + 512 + GUARD
    xin = (torch.randn((cin // 8, plen, 8), device="cuda") * 0.5).to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    w = (torch.randn((cout // nt, taps, cin // 8, nt, 8), device="cuda") * 0.05).to(torch.bfloat16)
    shift = torch.zeros(cout, device="cuda")
    out = torch.zeros((cout // 8, plen, 8), device="cuda", dtype=torch.bfloat16)
    mt = nacc * 128
    ncta = (P + mt - 1) // mt
    probe = torch.zeros((ncta, 16), device="cuda", dtype=torch.int64)
    flops = 2.0 * N * H * H * cin * cout * taps
    for dbg, label in ((0, "full"),):
        d = DebugConv(inp=xin[:, GUARD:].data_ptr(), in_plane_stride=plen * 8, w=w.data_ptr(), shift=shift.data_ptr(),
                      out=out[:, GUARD:].data_ptr(), out_plane_stride=plen * 8, res=None, res_plane_stride=0,
                      head_out=None, P=P, Hs=H + 1, Wp=H + 1, chunks=cin // 64, a_stages=2, act=1,
                      cout=cout, cout_pad=cout, nt=nt, nacc=nacc, taps=taps, impl=0, fmt=0, dbg=dbg, probe=None)
        for _ in range(2):
            assert lib.popnet_debug_conv(C.byref(d), None) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            lib.popnet_debug_conv(C.byref(d), None)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5 * 1e3
        print("%-18s %-18s %8.1f us  %7.1f TFLOP/s  (%d CTAs)" % (name, label, t, flops / t / 1e6, ncta))
    for dbg in (0, 2):
        d.dbg = dbg
        d.probe = probe.data_ptr()
        lib.popnet_debug_conv(C.byref(d), None)
        torch.cuda.synchronize()
        for ci in [int(v) for v in a.ctas.split(",")]:
            r = probe.cpu().numpy()[ci]
            nm = r[9] * nacc * 4 * taps * (cin // 64)
            print("   dbg %2d cta %3d: tiles %3d | mma thread: wait_a %7d wait_b %7d wait_acc %7d end +%7d (%.1f cyc/MMA net) | epilogue: "
              "wait %7d busy %7d | total %7d" % (dbg, ci, r[9], r[3], r[4], r[10], r[5] - r[1], (r[5] - r[1] - r[3] - r[4] - r[10]) / max(nm, 1),
                                     r[6], r[7], r[8] - r[0]))
