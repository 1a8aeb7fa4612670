"""Timeline of one forward (batch 64) from in-kernel %globaltimer stamps: for every conv / stem / pool launch the time
its first CTA started, the time its first CTA passed griddepcontrol.wait, and the time its last CTA finished.
Shows the real overlap of the three branch streams and the launch gaps that ncu's serialised list cannot."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popnet_b200 import network, synth, _abi, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--tuning", type=lambda v: int(v, 0), default=0, help="PopnetNetConfig.tuning bits (_abi.TUNE_*)")
a = ap.parse_args()
lib = _lib.get()
lib.popnet_debug_trace.restype = C.c_int
lib.popnet_debug_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]

m = network.rtpose_light3d(15, 14, 2, input_dim=1)
sd = network.synth_state_dict(seed=11, style="trained_like")
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m.tuning = a.tuning
x = torch.from_numpy(synth.depth_frames(8, seed=1)).cuda().repeat(a.batch // 8 + 1, 1, 1, 1)[:a.batch].contiguous()
for _ in range(5):
    m(x)
torch.cuda.synchronize()
CAP = 64
init = np.tile(np.array([2**64 - 1, 2**64 - 1, 0, 0], dtype=np.uint64), CAP)
buf = torch.from_numpy(init.view(np.int64)).cuda()
lib.popnet_debug_trace(C.c_void_p(buf.data_ptr()), CAP, None, 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m(x)
e1.record()
torch.cuda.synchronize()
tags = (C.c_int * CAP)()
n = lib.popnet_debug_trace(None, 0, tags, CAP)
r = buf.cpu().numpy().view(np.uint64).reshape(CAP, 4)[:n].astype(np.float64)
t0 = r[:, 0].min()
print("forward (events, traced): %.1f us; %d launches; span of stamps %.1f us" % (e0.elapsed_time(e1) * 1e3, n, (r[:, 2].max() - t0) / 1e3))
print("%3s %-14s %9s %9s %9s %8s %8s %9s" % ("#", "kernel", "start", "past-wait", "end", "active", "prologue", "SM-us/148"))
order = np.argsort(r[:, 0])
for i in order:
    t = tags[i]
    name = "stem" if t == 1 else "pool" if t == 2 else "conv<%d,%d,%d>%s" % (t // 1000, t // 100 % 10, t // 10 % 10, ("/chain%d" % (t % 10)) if t % 10 in (1, 2, 3, 4) and t // 10 % 10 == 9 and t // 1000 == 64 else "")
    print("%3d %-14s %9.1f %9.1f %9.1f %8.1f %8.1f %9.1f" % (i, name, (r[i, 0] - t0) / 1e3, (r[i, 1] - t0) / 1e3, (r[i, 2] - t0) / 1e3,
                                                         (r[i, 2] - r[i, 1]) / 1e3, (r[i, 1] - r[i, 0]) / 1e3, r[i, 3] / 1e3 / 148))
# SM time (sum over the CTAs of exit - entry, in units of the whole chip): what each launch occupies whatever queued before it
print("SM time / 148: total %.1f us of a %.1f us span (pools not counted)" % (r[:, 3].sum() / 1e3 / 148, (r[:, 2].max() - t0) / 1e3))
