#!/usr/bin/env python
"""Sensitivity of the end-to-end joint parity to 16-bit operand rounding, on the CPU (test infrastructure, not product code).

Runs the fp32 oracle forward and an EMULATION of popnet_forward's arithmetic (BatchNorm scale folded into the weights, weights
and every inter-layer activation rounded to fp16 / bf16, fp32 accumulation, fp32 heads -- csrc/forward.cu, csrc/conv_kernels.cu)
on the same synthetic frames, decodes both with the C oracle (bit-identical to the reference's paf_to_pose)
and classifies the frames whose assembled persons differ.  The emulation is not bit-identical to the device (summation order),
but it has the same rounding points, so it tells which checkpoints / decode decisions are marginal without spending GPU time.

    python tools/e2e_sensitivity.py [--ckpt tests/golden/fixture_ckpt.npz] [--frames 1024] [--fmt fp16]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from oracle import c_oracle, forward_torch  # noqa: E402
from popnet_b200 import _abi, synth  # noqa: E402
from popnet_b200.topology import MP3DHP, DecodeConfig  # noqa: E402


def rnd(x, fmt):
    return x.to(torch.float16 if fmt == "fp16" else torch.bfloat16).float()


def emulated_forward(sd, x, fmt, skip=()):
    """skip: rounding points left in fp32 (error attribution): any of "w", "block0", "stage1", "heads", "stage2"."""
    sd = {k: (torch.from_numpy(v) if not isinstance(v, torch.Tensor) else v) for k, v in sd.items()}
    fmt0 = fmt
    where = ["input"]

    def rnd(t, f, kind="act"):          # shadows the module-level rnd: honours `skip`
        if (kind == "w" and "w" in skip) or (kind == "act" and (where[0] in skip or ("block0" in skip and where[0] in
                                             ("input", "stem", "layer1", "pool1", "layer2", "conv2", "pool2")))):
            return t
        return t.to(torch.float16 if f == "fp16" else torch.bfloat16).float()

    def fold(prefix_conv, prefix_bn, bias=None):
        w = sd[prefix_conv + ".weight"]
        if prefix_bn is None:
            return rnd(w, fmt, "w"), (sd[bias] if bias else torch.zeros(w.shape[0]))
        g, b = sd[prefix_bn + ".weight"], sd[prefix_bn + ".bias"]
        m, v = sd[prefix_bn + ".running_mean"], sd[prefix_bn + ".running_var"]
        s = g / torch.sqrt(v + 1e-5)
        shift = b - m * s
        if bias:
            shift = shift + sd[bias] * s
        return rnd(w * s[:, None, None, None], fmt, "w"), shift

    def conv(x, w, shift, stride=1):
        return F.conv2d(x, w, None, stride, w.shape[2] // 2) + shift[None, :, None, None]

    w, s = fold("model0.conv1", "model0.bn1")
    xin = rnd(x, fmt)
    if where[0] not in skip and "block0" not in skip:
        xin = xin + rnd(x - xin, fmt)          # the stem multiplies hi = rn16(x) and lo = rn16(x - hi) (conv_kernels.cu, stem)
    where[0] = "stem"
    y = rnd(F.relu(conv(xin, w, s, 2)), fmt)

    def block(y, p, down):
        w1, s1 = fold(p + ".conv1", p + ".bn1")
        w2, s2 = fold(p + ".conv2", p + ".bn2")
        o = rnd(F.relu(conv(y, w1, s1)), fmt)
        o = conv(o, w2, s2)
        if down:
            wd, sdn = fold(p + ".downsample.0", p + ".downsample.1")
            o = o + conv(y, wd, sdn)
        else:
            o = o + y
        return rnd(F.relu(o), fmt)

    where[0] = "layer1"
    y = block(y, "model0.layer1.0", False)
    y = block(y, "model0.layer1.1", False)
    where[0] = "pool1"
    y = rnd(F.avg_pool2d(y, 3, 2, 1), fmt)
    where[0] = "layer2"
    y = block(y, "model0.layer2.0", True)
    w, s = fold("model0.conv2", "model0.bn2")
    where[0] = "conv2"
    y = rnd(F.relu(conv(y, w, s)), fmt)
    where[0] = "pool2"
    out1 = rnd(F.avg_pool2d(y, 3, 2, 1), fmt)

    def stage(x, name):
        for i in range(5):
            cn = "%s.%d" % (name, 3 * i)
            if i < 4:
                w, s = fold(cn, "%s.%d" % (name, 3 * i + 1), cn + ".bias")
                x = rnd(F.leaky_relu(conv(x, w, s), 0.1), fmt)
            else:
                w, s = fold(cn, None, cn + ".bias")
                x = conv(x, w, s)
        return x

    where[0] = "stage1"
    paf1 = (stage(out1, "model1_1").sigmoid() - 0.5) * 4
    heat1 = stage(out1, "model1_2").sigmoid()
    dep1 = (stage(out1, "model1_3").sigmoid() - 0.5) * 4
    where[0] = "heads"
    out2 = torch.cat([rnd(paf1, fmt), rnd(heat1, fmt), rnd(dep1, fmt), out1], 1)
    where[0] = "stage2"
    paf2 = (stage(out2, "model2_1").sigmoid() - 0.5) * 4
    heat2 = stage(out2, "model2_2").sigmoid()
    dep2 = (stage(out2, "model2_3").sigmoid() - 0.5) * 4
    return paf2, heat2, dep2


def compare(a, b, f):
    """frame f of two record dicts -> 'exact' | 'structural' | reason"""
    na, nb = int(a["n_person"][f]), int(b["n_person"][f])
    if na != nb:
        return "n_person %d vs %d" % (na, nb)
    pa, pb = a["pose2d"][f, :na, :15], b["pose2d"][f, :nb, :15]
    if np.array_equal(pa, pb):
        return "exact"
    va, vb = pa[:, :, 0] >= 0, pb[:, :, 0] >= 0
    if not np.array_equal(va, vb):
        return "visible-set (%d joints differ)" % int((va != vb).sum())
    d = np.abs(pa - pb)[va]
    if (d[:, 0] <= 480.0 / 224 + 1e-9).all() and (d[:, 1] <= 512.0 / 224 + 1e-9).all():
        return "structural"
    return "coords (max %.1f px)" % float(d.max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt", default=os.path.join(ROOT, "tests", "golden", "fixture_ckpt.npz"))
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=777_000)
    ap.add_argument("--fmt", default="fp16")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--skip", default="", help="comma-separated rounding points kept in fp32 (error attribution): w,block0 (= input,stem,layer1,pool1,layer2,conv2,pool2),stage1,heads,stage2")
    args = ap.parse_args()
    torch.set_num_threads(args.threads or os.cpu_count() or 1)
    z = np.load(args.ckpt)
    sd = {k: (z[k].astype(np.float32) if z[k].dtype != np.int64 else z[k]) for k in z.files}
    x = synth.depth_frames(args.frames, seed=args.seed)
    params = _abi.make_decode_params(DecodeConfig(), MP3DHP, max_persons=32)
    recs = {"ref": [], "emu": []}
    err = np.zeros(3)
    with torch.no_grad():
        for b0 in range(0, args.frames, 32):
            xb = torch.from_numpy(x[b0:b0 + 32])
            (paf, heat, depth), _ = forward_torch.forward(sd, xb)
            e = emulated_forward(sd, xb, args.fmt, tuple(v for v in args.skip.split(",") if v))
            for i, (u, v) in enumerate(zip((paf, heat, depth), e)):
                err[i] = max(err[i], float((u - v).abs().max()))
            recs["ref"].append(c_oracle.decode(heat.numpy(), paf.numpy(), depth.numpy(), params))
            recs["emu"].append(c_oracle.decode(e[1].numpy(), e[0].numpy(), e[2].numpy(), params))
            print("frames", b0 + 32, flush=True)
    cat = {k: {n: np.concatenate([r[n] for r in v], 0) for n in v[0]} for k, v in recs.items()}
    res = {}
    bad = []
    for f in range(args.frames):
        c = compare(cat["ref"], cat["emu"], f)
        key = c.split(" ")[0]
        res[key] = res.get(key, 0) + 1
        if c not in ("exact", "structural"):
            bad.append((f, c))
    print("max-abs error paf/heat/depth:", err)
    print("persons ref/emu:", int(cat["ref"]["n_person"].sum()), int(cat["emu"]["n_person"].sum()),
          "flags:", int((cat["ref"]["flags"] != 0).sum()), int((cat["emu"]["flags"] != 0).sum()))
    print(res)
    ok = res.get("exact", 0) + res.get("structural", 0)
    print("structural-or-better: %d / %d = %.4f" % (ok, args.frames, ok / args.frames))
    for f, c in bad[:40]:
        print("  frame", f, c)


if __name__ == "__main__":
    main()
