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
a = ap.parse_args()

m = network.rtpose_light3d(15, 14, 2, input_dim=1)
sd = network.synth_state_dict(seed=11, style="trained_like")
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m.operand_dtype = _abi.OPERAND_BF16 if a.dtype == "bf16" else _abi.OPERAND_FP16
m.impl = _abi.FWD_IMPL_TCGEN05 if a.impl == "tc" else _abi.FWD_IMPL_SIMT
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
print("forward batch %d: median %.3f ms  (%.0f frames/s, %.1f TFLOP/s)  all: %s" %
      (a.batch, t, a.batch / t * 1e3, a.batch * 13.343404032 / t, ["%.3f" % v for v in ts]))
