#!/usr/bin/env python
"""Continue training the FIXTURE checkpoint on a GPU box (test infrastructure, not product code).

tools/make_fixture_ckpt.py trains the UNMODIFIED reference module with the reference's own loss, on the CPU of the build
container -- a few thousand steps of batch 8 is all that fits there, and the resulting maps are still blobby (persons come
out fragmented, a percent of the frames hold a decode decision within 1e-3 of its threshold).  /root/reference does not exist
on the GPU box, so this script continues from that checkpoint with a train-mode restatement of the same module (the
functional oracle forward, oracle/forward_torch.py, with batch-statistics BatchNorm: rtpose_light3d.py:124-356) and of the
same loss (lib/network/losses.py:65-106: MSE on the PAF and heat maps, foreground-weighted MSE (0.1 + 0.9 fg) on the depth
maps, summed over both stages), for a wall-clock budget.  The parameterisation is the reference's 234-key state dict, so the
result loads into the reference module unchanged; tests/golden/make_golden.py e2e then regenerates the end-to-end golden
with the LIVE reference on that checkpoint (build container).

    python tools/train_fixture_gpu.py --seconds 600 --out gpurun_out/fixture_gpu.npz
"""
import argparse
import math
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

from popnet_b200 import synth  # noqa: E402
from popnet_b200.topology import MP3DHP  # noqa: E402


def _pool_chunk(args):
    seed, n, persons = args
    x = synth.depth_frames(n, seed=seed, persons=persons)
    heat, paf, depth, _ = synth.map_batch(n, seed=seed, persons=persons, noise=0.0)
    bg = np.float32((MP3DHP.depth_max - MP3DHP.depth_mean) / MP3DHP.depth_std)
    fg = (depth != bg).astype(np.float32)
    return x, heat, paf, depth, fg


def make_pool(n, seed, persons, procs):
    chunk = 64
    jobs = [(seed + i, min(chunk, n - i), persons) for i in range(0, n, chunk)]
    with mp.Pool(procs) as pool:
        parts = pool.map(_pool_chunk, jobs)
    return [np.concatenate([p[i] for p in parts], 0) for i in range(5)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=600.0, help="training wall-clock budget")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--pool", type=int, default=16384)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=100_000, help="frame seeds [seed, seed+pool): disjoint from every test / bench seed")
    ap.add_argument("--resume", default=os.path.join(ROOT, "tests", "golden", "fixture_ckpt.npz"))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fixture_gpu.npz"))
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()

    import torch
    import torch.nn.functional as F
    dev = args.device
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True

    t0 = time.time()
    x, heat, paf, depth, fg = (torch.from_numpy(a).to(dev) for a in make_pool(args.pool, args.seed, (1, 6), os.cpu_count() or 8))
    print("pool of %d frames in %.0f s" % (args.pool, time.time() - t0), flush=True)

    z = np.load(args.resume)
    sd = {}
    for k in z.files:
        v = z[k]
        t = torch.from_numpy(v.astype(np.float32) if v.dtype != np.int64 else v).to(dev)
        sd[k] = t
    train_keys = [k for k in sd if sd[k].dtype == torch.float32 and not (k.endswith("running_mean") or k.endswith("running_var"))]
    for k in train_keys:
        sd[k].requires_grad_(True)

    def bn(y, prefix):
        return F.batch_norm(y, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"], sd[prefix + ".bias"],
                            True, 0.1, 1e-5)

    def block(y, prefix, has_down):
        out = F.relu(bn(F.conv2d(y, sd[prefix + ".conv1.weight"], None, 1, 1), prefix + ".bn1"))
        out = bn(F.conv2d(out, sd[prefix + ".conv2.weight"], None, 1, 1), prefix + ".bn2")
        idt = y
        if has_down:
            idt = bn(F.conv2d(y, sd[prefix + ".downsample.0.weight"]), prefix + ".downsample.1")
        return F.relu(out + idt)

    def stage(y, name):
        for i in range(5):
            w = sd["%s.%d.weight" % (name, 3 * i)]
            y = F.conv2d(y, w, sd["%s.%d.bias" % (name, 3 * i)], 1, w.shape[2] // 2)
            if i < 4:
                y = F.leaky_relu(bn(y, "%s.%d" % (name, 3 * i + 1)), 0.1)
        return y

    def forward(xb):
        y = F.relu(bn(F.conv2d(xb, sd["model0.conv1.weight"], None, 2, 3), "model0.bn1"))
        y = block(y, "model0.layer1.0", False)
        y = block(y, "model0.layer1.1", False)
        y = F.avg_pool2d(y, 3, 2, 1)
        y = block(y, "model0.layer2.0", True)
        y = F.relu(bn(F.conv2d(y, sd["model0.conv2.weight"]), "model0.bn2"))
        out1 = F.avg_pool2d(y, 3, 2, 1)
        paf1 = (stage(out1, "model1_1").sigmoid() - 0.5) * 4
        heat1 = stage(out1, "model1_2").sigmoid()
        dep1 = (stage(out1, "model1_3").sigmoid() - 0.5) * 4
        out2 = torch.cat([paf1, heat1, dep1, out1], 1)
        paf2 = (stage(out2, "model2_1").sigmoid() - 0.5) * 4
        heat2 = stage(out2, "model2_2").sigmoid()
        dep2 = (stage(out2, "model2_3").sigmoid() - 0.5) * 4
        return [paf1, heat1, dep1, paf2, heat2, dep2]

    def loss_fn(saved, ib):
        # the third head has num_limbs + 1 = 15 planes = one per joint (rtpose_light3d.py:299-309)
        w = 0.1 + 0.9 * fg[ib]
        total, parts = 0.0, []
        for j in range(2):
            l1 = F.mse_loss(saved[3 * j], paf[ib])
            l2 = F.mse_loss(saved[3 * j + 1], heat[ib])
            l3 = (w * (saved[3 * j + 2] - depth[ib]) ** 2).mean()
            total = total + l1 + l2 + l3
            parts += [l1, l2, l3]
        return total, parts

    opt = torch.optim.Adam([sd[k] for k in train_keys], lr=args.lr)
    g = torch.Generator(device=dev).manual_seed(1)

    def save(path):
        out = {}
        for k, v in sd.items():
            a = v.detach().cpu().numpy()
            if a.dtype == np.float32 and not (k.endswith("running_mean") or k.endswith("running_var")):
                a = a.astype(np.float16)
            out[k] = a
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **out)

    t0 = time.time()
    step = 0
    ema = None
    while True:
        el = time.time() - t0
        if el >= args.seconds:
            break
        frac = el / args.seconds
        lr = args.lr * min(1.0, (step + 1) / 200.0) * (0.02 + 0.98 * 0.5 * (1.0 + math.cos(math.pi * frac)))
        for pg in opt.param_groups:
            pg["lr"] = lr
        ib = torch.randint(0, args.pool, (args.batch,), generator=g, device=dev)
        saved = forward(x[ib])
        total, parts = loss_fn(saved, ib)
        opt.zero_grad(set_to_none=True)
        total.backward()
        opt.step()
        step += 1
        if step % 100 == 0:
            v = float(total)
            ema = v if ema is None else 0.9 * ema + 0.1 * v
            print("step %6d  %.0f s  lr %.2e  loss %.5f (ema %.5f)  %s" % (
                step, el, lr, v, ema, " ".join("%.4f" % float(p) for p in parts)), flush=True)
        if step % 4000 == 0:
            save(args.out)
    for k in sd:
        if k.endswith("num_batches_tracked"):
            sd[k] += step
    save(args.out)
    print("wrote %s after %d steps, %.1f MB" % (args.out, step, os.path.getsize(args.out) / 1e6))


if __name__ == "__main__":
    main()
