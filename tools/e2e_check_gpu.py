#!/usr/bin/env python
"""End-to-end joint parity of a checkpoint on the GPU box, without the golden (test infrastructure, not product code).

Path A: fp32 oracle forward (cuDNN, TF32 off) -> C-oracle decode (bit-identical to the reference's paf_to_pose).
Path B: popnet_b200.pipeline.PoseEstimator (16-bit operands) on the same frames.
Control: the oracle forward with TF32 allowed -- what the reference's own GPU path runs by default (torch's
cudnn.allow_tf32 = True) -- decoded by the same C oracle: how often the REFERENCE disagrees with itself at 10 mantissa bits.

    python tools/e2e_check_gpu.py --ckpt gpurun_out/fixture_gpu.npz --out gpurun_out/e2e_check.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]

from e2e_sensitivity import compare  # noqa: E402
from oracle import c_oracle, forward_torch  # noqa: E402
from popnet_b200 import _abi, network, pipeline, synth  # noqa: E402


def oracle_records(sd, x, params, tf32):
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    recs, maps = [], []
    try:
        for b0 in range(0, len(x), 64):
            (paf, heat, depth), _ = forward_torch.forward(sd, torch.from_numpy(x[b0:b0 + 64]).cuda())
            paf, heat, depth = paf.cpu().numpy(), heat.cpu().numpy(), depth.cpu().numpy()
            recs.append(c_oracle.decode(heat, paf, depth, params))
            maps.append((paf, heat, depth))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    return {k: np.concatenate([r[k] for r in recs], 0) for k in recs[0]}, maps


def tally(a, b, n):
    res, bad = {}, []
    for f in range(n):
        c = compare(a, b, f)
        res[c.split(" ")[0]] = res.get(c.split(" ")[0], 0) + 1
        if c not in ("exact", "structural"):
            bad.append((f, c))
    res["ok"] = res.get("exact", 0) + res.get("structural", 0)
    return res, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt", default=os.path.join(ROOT, "tests", "golden", "fixture_ckpt.npz"))
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=777_000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    z = np.load(args.ckpt)
    sd = {k: (z[k].astype(np.float32) if z[k].dtype != np.int64 else z[k]) for k in z.files}
    x = synth.depth_frames(args.frames, seed=args.seed)
    m = network.rtpose_light3d(15, 14, 2, input_dim=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    out = {"ckpt": os.path.basename(args.ckpt), "frames": args.frames}
    ref = None
    for dtype in ("fp16", "bf16"):
        m.operand_dtype = _abi.OPERAND_FP16 if dtype == "fp16" else _abi.OPERAND_BF16
        est = pipeline.PoseEstimator(m, max_persons=32, strict=False)
        if ref is None:
            ref, ref_maps = oracle_records(sd, x, est.params, False)
            ctl, _ = oracle_records(sd, x, est.params, True)
            r, bad = tally(ref, ctl, args.frames)
            out["control_tf32_vs_fp32"] = r
            print("control (reference fp32 vs reference TF32):", r, bad[:8], flush=True)
            out["persons_reference"] = int(ref["n_person"].sum())
        recs, err = [], np.zeros(3)
        for b0 in range(0, args.frames, 64):
            recs.append({k: np.array(v) for k, v in est.infer(x[b0:b0 + 64]).items()})
            (paf, heat, depth), _ = m(torch.from_numpy(x[b0:b0 + 64]).cuda())
            for i, (u, v) in enumerate(zip((paf, heat, depth), ref_maps[b0 // 64])):
                err[i] = max(err[i], float(np.abs(u.cpu().numpy() - v).max()))
        rec = {k: np.concatenate([r[k] for r in recs], 0) for k in recs[0]}
        r, bad = tally(ref, rec, args.frames)
        out[dtype] = {"parity": r, "max_abs_paf_heat_depth": [float(e) for e in err], "bad": bad[:16],
                      "flags": int((rec["flags"] != 0).sum())}
        print(dtype, out[dtype], flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f)


if __name__ == "__main__":
    main()
