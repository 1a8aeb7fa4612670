"""Quick device timing of the forward + decode at a given batch (CUDA events, L2-flushing input rotation)."""
import argparse
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popnet_b200 import network, synth, _abi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--impl", default="tc")
ap.add_argument("--streams", type=int, default=1)
ap.add_argument("--tuning", type=lambda v: int(v, 0), default=0, help="PopnetNetConfig.tuning bits (_abi.TUNE_*)")
a = ap.parse_args()

m = network.rtpose_light3d(15, 14, 2, input_dim=1)
sd = network.synth_state_dict(seed=11, style="trained_like")
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m.operand_dtype = _abi.OPERAND_BF16 if a.dtype == "bf16" else _abi.OPERAND_FP16
m.impl = _abi.FWD_IMPL_TCGEN05 if a.impl == "tc" else _abi.FWD_IMPL_SIMT
m.tuning = a.tuning
x = torch.from_numpy(synth.depth_frames(8, seed=1)).cuda().repeat(a.batch // 8 + 1, 1, 1, 1)[:a.batch].contiguous()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
ev[0].record()
for i in range(a.iters):
    m(x)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]
t = float(np.median(ts))
print("tuning 0x%x" % a.tuning, end=" ")
print("forward batch %d: median %.3f ms  (%.0f frames/s, %.1f TFLOP/s)  all: %s" %
      (a.batch, t, a.batch / t * 1e3, a.batch * 13.343404032 / t, ["%.3f" % v for v in ts]))

# ---- optional: K forwards in flight on K streams (K model instances = K workspaces); aggregate throughput
if a.streams > 1:
    ms = [m]
    for _ in range(a.streams - 1):
        mm = network.rtpose_light3d(15, 14, 2, input_dim=1)
        mm.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        mm.operand_dtype, mm.impl = m.operand_dtype, m.impl
        ms.append(mm)
    sts = [torch.cuda.Stream() for _ in ms]
    xs = [x.clone() for _ in ms]
    torch.cuda.synchronize()

    def run(n):
        for i in range(n):
            k = i % len(ms)
            with torch.cuda.stream(sts[k]):
                ms[k](xs[k])
    run(2 * len(ms))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in sts:
        s_.wait_event(e0)
    run(a.iters)
    for s_ in sts:
        torch.cuda.current_stream().wait_stream(s_)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / a.iters
    print("%d streams: %.3f ms per forward of %d  (%.0f frames/s, %.1f TFLOP/s)" %
          (a.streams, t, a.batch, a.batch / t * 1e3, a.batch * 13.343404032 / t))
