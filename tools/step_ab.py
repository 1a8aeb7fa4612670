"""A/B of the step schedule in sustained operation (>= 2 s per variant, CUDA events): forward only, forward + decode
overlapped on the decode stream (the product schedule), forward + decode serialised on one stream.  Tuning aid."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from popnet_b200 import _abi, network, pipeline, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--tuning", type=lambda v: int(v, 0), default=0)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--decode-priority", type=int, default=0)
ap.add_argument("--decode-ctas", type=int, default=-1, help="CTA limit of the decode kernels (-1 = the reserve, 0 = all SMs)")
ap.add_argument("--reserve", type=int, default=8, help="SMs the conv grids leave to the decode (0 = none)")
ap.add_argument("--only", default="", help="comma-separated variant names to run (default: all)")
a = ap.parse_args()

z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fixture_ckpt.npz"))
sd = {k: (z[k].astype(np.float32) if z[k].dtype != np.int64 else z[k]) for k in z.files}
m = network.rtpose_light3d(15, 14, 2, input_dim=1)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m.operand_dtype = _abi.OPERAND_FP16
m.tuning = a.tuning
est = pipeline.PoseEstimator(m, max_persons=32, strict=False, decode_priority=a.decode_priority, reserve_sms=a.reserve, decode_ctas=None if a.decode_ctas < 0 else a.decode_ctas)
B = a.batch
NS = est.NSLOT
for i in range(NS):
    est.slot_input(i, B).copy_(torch.from_numpy(synth.depth_frames(B, seed=1000 + 100 * i)))
for i in range(2 * NS):
    est.infer_device(est.slot_input(i, B))
torch.cuda.synchronize()
slots = est._slots
main, ds = torch.cuda.current_stream(), est.decode_stream


def fwd_only(i):
    slots[i % NS]["graphs"][0].replay()


def overlapped(i):
    est.infer_device(est.slot_input(i, B))


def serial(i):
    s = slots[i % NS]
    s["graphs"][0].replay()
    s["graphs"][1].replay()


def dec_only(i):
    slots[i % NS]["graphs"][1].replay()


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    main.wait_stream(ds)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for r in range(a.rounds):
    for name, fn in (("forward only", fwd_only), ("overlapped (product)", overlapped), ("serialised", serial), ("decode only", dec_only)):
        if a.only and name.split(" ")[0] not in a.only.split(","):
            continue
        t = timed(fn, 50)
        n = max(50, int(a.seconds * 1e3 / t))
        t = timed(fn, n)
        print("tuning 0x%x reserve %d decode ctas %d: %-22s %.4f ms per step  (%6.0f frames/s)  [%d steps]" % (a.tuning, a.reserve, a.decode_ctas, name, t, B / t * 1e3, n), flush=True)
