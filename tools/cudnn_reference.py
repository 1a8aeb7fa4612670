"""GPU reference bar: the same rtpose_light3d architecture executed by PyTorch eager / cuDNN on the
B200 -- fp32, TF32 and bf16 + channels_last -- timed at batch 64 next to popnet_forward.  The eager forward lives
HERE (a measurement tool); the product has no torch path."""
import argparse
import copy
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popnet_b200 import network, synth  # noqa: E402


def eager_forward(m, x):
    """rtpose_light3d.py:201-216,326-356 with torch ops on the parameter containers."""
    m0 = m.model0
    y = F.relu(m0.bn1(m0.conv1(x)))
    for blk in m0.layer1:
        y = F.relu(blk.bn2(blk.conv2(F.relu(blk.bn1(blk.conv1(y))))) + y)
    y = F.avg_pool2d(y, 3, 2, 1)
    blk = m0.layer2[0]
    y = F.relu(blk.bn2(blk.conv2(F.relu(blk.bn1(blk.conv1(y))))) + blk.downsample(y))
    y = F.relu(m0.bn2(m0.conv2(y)))
    feat = F.avg_pool2d(y, 3, 2, 1)
    outs = []
    inp = feat
    for s in (1, 2):
        paf = (torch.sigmoid(getattr(m, "model%d_1" % s)(inp)) - 0.5) * 4
        heat = torch.sigmoid(getattr(m, "model%d_2" % s)(inp))
        dep = (torch.sigmoid(getattr(m, "model%d_3" % s)(inp)) - 0.5) * 4
        outs.append((paf, heat, dep))
        inp = torch.cat([paf, heat, dep, feat], 1)
    return outs[-1]


def time_it(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
torch.backends.cudnn.benchmark = True
m = network.rtpose_light3d(15, 14, 2, input_dim=1)
sd = network.synth_state_dict(seed=0, style="reference")
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
x = torch.from_numpy(synth.depth_frames(8, seed=1)).cuda().repeat(a.batch // 8 + 1, 1, 1, 1)[:a.batch].contiguous()
res = {"batch": a.batch, "flop_per_frame": 13343404032}
ours = time_it(lambda: m(x), a.iters * 3)
res["popnet_forward_bf16_ms"] = ours
(p0, h0, d0), _ = m(x)
ref = copy.deepcopy(m).cuda().eval()
with torch.no_grad():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res["cudnn_fp32_ms"] = time_it(lambda: eager_forward(ref, x), a.iters)
    pr, hr, dr = eager_forward(ref, x)
    res["max_abs_diff_vs_cudnn_fp32"] = float(max((p0 - pr).abs().max(), (h0 - hr).abs().max(), (d0 - dr).abs().max()))
    torch.backends.cudnn.allow_tf32 = True
    res["cudnn_tf32_ms"] = time_it(lambda: eager_forward(ref, x), a.iters)
    refb = copy.deepcopy(ref).to(torch.bfloat16).to(memory_format=torch.channels_last)
    xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    res["cudnn_bf16_channels_last_ms"] = time_it(lambda: eager_forward(refb, xb), a.iters)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        res["cudnn_autocast_bf16_ms"] = time_it(lambda: eager_forward(ref, x), a.iters)
for k in list(res):
    if k.endswith("_ms"):
        res[k.replace("_ms", "_tflops")] = a.batch * 13.343404032 / res[k]
print(json.dumps(res))
