#!/usr/bin/env python
"""Train the FIXTURE checkpoint: the UNMODIFIED reference module (lib/network/rtpose_light3d.py) with the
reference's own loss (lib/network/losses.py:65-106, rtpose_light3d_loss_fgweight) and the reference's training-step
body (train_rtpose_light3d_kdh3d_mpaug.py:153-213: train mode, forward, loss over the six saved maps, zero_grad,
backward, step) on synthetic depth frames and the GT maps rendered from the SAME skeletons
(popnet_b200.synth.depth_frames / map_batch share their per-frame seeds).

No trained weights ship with the reference (README.md:43-45: trained_model/ is in the 800 GB torrent only), and
reference-initialised weights give sigma ~ 0.5 heat-maps that cannot be decoded ("Hard parts").
This script is the fixture generator SURVEY.md section 7 (iv) describes: test infrastructure, not product code.
It runs in the BUILD container (where /root/reference exists), on the CPU:

    python tools/make_fixture_ckpt.py --steps 1500 --batch 8          # ~1 s per step on 8 cores

and writes tests/golden/fixture_ckpt.npz: the 234 state-dict tensors, fp16 (BatchNorm statistics fp32), compressed.
Loading it (tests/helpers.fixture_state_dict) widens back to fp32: the fixture IS the fp16-rounded checkpoint, used
identically by the reference, the oracle and the CUDA path.

Deviations from the reference's training script, all outside the parity surface: Adam instead of SGD(lr=1, m=0.9)
(converges in hundreds of steps instead of epochs), no DataParallel, CPU instead of .cuda(), a fixed frame pool
instead of the mp-aug dataset.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]

import refshim  # noqa: E402
from popnet_b200 import synth  # noqa: E402
from popnet_b200.topology import MP3DHP  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fixture_ckpt.npz")


def make_pool(n, seed, persons):
    """frames [n,1,224,224], heat [n,16,28,28], paf [n,28,28,28], depth [n,15,28,28], fg mask [n,15,28,28]."""
    x = synth.depth_frames(n, seed=seed, persons=persons)
    heat, paf, depth, _ = synth.map_batch(n, seed=seed, persons=persons, noise=0.0)
    bg = np.float32((MP3DHP.depth_max - MP3DHP.depth_mean) / MP3DHP.depth_std)
    fg = (depth != bg).astype(np.float32)          # the joint patches of posemap.py:83-106
    return x, heat, paf, depth, fg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--pool", type=int, default=1536)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=50_000, help="frame seeds [seed, seed+pool): disjoint from every test / bench seed")
    ap.add_argument("--resume", default=None)
    ap.add_argument("--out", default=OUT)
    ap.add_argument("--save-every", type=int, default=250)
    ap.add_argument("--threads", type=int, default=0, help="torch CPU threads (0 = all cores)")
    args = ap.parse_args()

    import torch
    torch.manual_seed(0)
    torch.set_num_threads(args.threads or os.cpu_count() or 1)
    ref = refshim.load()
    from lib.network.losses import rtpose_light3d_loss_fgweight          # the reference's loss, unmodified
    names = ["loss_stage%d_L%d" % (j, k) for j in (1, 2) for k in (1, 2, 3)]

    t0 = time.time()
    x, heat, paf, depth, fg = (torch.from_numpy(a) for a in make_pool(args.pool, args.seed, (1, 6)))
    print("pool of %d frames in %.0f s" % (args.pool, time.time() - t0), flush=True)

    model = ref.rtpose_light3d(15, 14, 2, input_dim=1).float()
    if args.resume:
        sd = {k: torch.from_numpy(v.astype(np.float32)) if v.dtype != np.int64 else torch.from_numpy(v)
              for k, v in np.load(args.resume).items()}
        model.load_state_dict(sd)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=args.lr)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=args.lr, total_steps=args.steps, pct_start=0.1)
    g = torch.Generator().manual_seed(1)

    def save():
        out = {}
        for k, v in model.state_dict().items():
            a = v.detach().cpu().numpy()
            if a.dtype == np.float32 and not (k.endswith("running_mean") or k.endswith("running_var")):
                a = a.astype(np.float16)
            out[k] = a
        np.savez_compressed(args.out, **out)

    model.train()
    t0 = time.time()
    for step in range(args.steps):
        idx = torch.randint(0, args.pool, (args.batch,), generator=g)
        _, saved_for_loss = model(x[idx])
        total_loss, log = rtpose_light3d_loss_fgweight(saved_for_loss, heat[idx], paf[idx], depth[idx], fg[idx], 2, names)
        opt.zero_grad()
        total_loss.backward()
        opt.step()
        sched.step()
        if step % 10 == 0 or step == args.steps - 1:
            print("step %5d  loss %.5f  %s  max_ht %.3f  %.1f s" % (
                step, float(total_loss), " ".join("%.4f" % log[n] for n in names), log["max_ht"], time.time() - t0), flush=True)
        if (step + 1) % args.save_every == 0 or step == args.steps - 1:
            save()
    print("wrote", args.out, "%.1f MB" % (os.path.getsize(args.out) / 1e6))


if __name__ == "__main__":
    main()
