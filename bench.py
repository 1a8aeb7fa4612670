#!/usr/bin/env python
"""Headline benchmark: depth frames/sec for rtpose_light3d forward + PAF decode + 3D lift (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dtype bf16|fp16]

N > 1 is launched by the driver through torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
A "step" is one pass of the hot path over one batch of synthetic depth frames per GPU:
    forward (39 convolutions on the tensor cores) -> decode (peaks, limbs, assembly) -> 3D lift
    [-> all-gather of the pose records when N > 1].
`value`   : inputs already resident in HBM, timed with CUDA events, max over ranks.
`e2e`     : the same step through the public API (popnet_b200.pipeline.PoseEstimator.infer) with pinned HOST
            frames in and pose records copied back to the host every step.
`--impl reference` times the CPU restatement of the reference's path (oracle/: fp32 torch forward + C decode)
on all host cores on a bounded sample of the same workload -- the reference itself is Python and cannot
travel to the GPU box (/root/reference is not there).
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FRAME = 13_343_404_032          # 2 x 6,671,702,016 conv MACs at 224x224, K=15, L=14
DECODE_BYTES_PER_FRAME = 4 * 784 * (15 + 28 + 15)
METRIC = "depth frames/sec (fwd+PAF decode+3D lift)"
WORKLOAD = ("C2: rtpose_light3d inference, batch 64 synthetic 224x224 depth frames per GPU, heatmaps+PAF+depth maps "
            "and greedy assembly + 3D lift")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", 1400.0), "tflops_burst": d.get("bf16_tflops", 1590.0),
                "hbm": d.get("hbm_gbs", 6650.0), "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(mhz)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_path(sample_frames, repeats, warmup, cores):
    """Returns (frames/s, seconds per repeat list).  forward: fp32 torch oracle with `cores` threads;
    decode + lift: C oracle, frames spread over `cores` threads (ctypes releases the GIL)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle, forward_torch
    from popnet_b200 import _abi, network, synth
    from popnet_b200.topology import MP3DHP, DecodeConfig
    torch.set_num_threads(cores)
    sd = network.synth_state_dict(seed=0, style="reference")
    x = torch.from_numpy(synth.depth_frames(sample_frames, seed=1234))
    heat, paf, depth, _ = synth.map_batch(sample_frames, seed=1234, persons=(1, 6), noise=0.01)
    params = _abi.make_decode_params(DecodeConfig(), MP3DHP, max_persons=32)
    c_oracle.lib()
    pool = ThreadPoolExecutor(cores)

    def one():
        forward_torch.forward(sd, x)
        list(pool.map(lambda f: c_oracle.decode(heat[f:f + 1], paf[f:f + 1], depth[f:f + 1], params), range(sample_frames)))

    for _ in range(warmup):
        one()
    ts = []
    for _ in range(repeats):
        t = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t)
    return sample_frames / float(np.mean(ts)), ts


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 16
    steps = min(args.steps, 40)          # bounded: the whole run ends within a few minutes
    fps, ts = cpu_path(sample, steps, min(args.warmup, 3), cores)
    desc = "%d frames per step: fp32 torch forward (%d threads) + C-oracle decode/lift over %d threads" % (sample, cores, cores)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": float(np.mean(ts)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU port of the reference path (oracle/); the Python reference "
                       "cannot travel to the GPU box"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def evaluator_leg(n_frames=4000, iters=20, with_cpu=True):
    """Secondary metric of SURVEY.md 8(d): the best-match evaluator (3D PCK matching + 3D mAP assignment) on the C3 set.
    `device`: kernels only, CSR arrays resident in HBM (CUDA events); `e2e`: the public reference-signature calls on
    the ragged Python lists (host packing, H2D, kernels, D2H, AP tail); `cpu_port`: the C oracle on the same arrays."""
    import io
    import contextlib
    import torch
    from popnet_b200 import evaluate, synth, topology
    from popnet_b200._cuda_backend import CudaBackend, _to_dev
    K = 15
    es = synth.eval_set(n_frames, seed=0)
    names = list(topology.JOINT_NAMES)
    captured = {}
    be = CudaBackend()

    class Tap:                                     # records the packed CSR arrays the public API hands to the backend
        def pck(self, arrs, **kw):
            captured["pck"] = (arrs, kw)
            return be.pck(arrs, **kw)

        def map_assign(self, arrs, **kw):
            captured["map"] = (arrs, kw)
            return be.map_assign(arrs, **kw)

    evaluate._backend = Tap()
    sink = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        evaluate.eval_human_dataset_3d(es["pred2d"], es["gt2d"], es["pred3d"], es["gt3d"], K, 0.1, 0.5)
        evaluate.eval_ap_3D(es["pred3d"], es["conf"], es["gt3d"], [], names, 0.1)
    e2e_s = time.perf_counter() - t0
    evaluate._backend = None
    (pa, pkw), (ma, mkw) = captured["pck"], captured["map"]
    dp = {k: _to_dev(v) for k, v in pa.items()}
    dm = {k: _to_dev(v) for k, v in ma.items()}
    op = be.pck_device(dp, **pkw)
    om = be.map_assign_device(dm, **mkw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        be.pck_device(dp, out=op, **pkw)
        be.map_assign_device(dm, out=om, **mkw)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / iters
    n_pred, n_gt = int(pa["pred2d"].shape[0]), int(pa["gt2d"].shape[0])
    alg_bytes = 8 * K * (n_pred * (2 + 3 + 1) + n_gt * (2 + 3)) + 8 * K * n_gt          # reads + distances written
    res = {"workload": "C3: %d frames, %d GT / %d predicted humans; eval_human_dataset_3d + eval_ap_3D" % (n_frames, n_gt, n_pred),
           "device": {"value": n_frames / (dev_ms * 1e-3), "unit": "frames/s", "ms": dev_ms, "launches": 2,
                      "achieved_GBps": alg_bytes / (dev_ms * 1e-3) / 1e9, "algorithmic_bytes": alg_bytes,
                      "bound": "hbm (latency-bound: one warp per frame; the 19 MB of CSR arrays stay in the 126 MB L2 between iterations)"},
           "e2e": {"value": n_frames / e2e_s, "unit": "frames/s", "s": e2e_s,
                   "note": "public API on ragged Python lists: list->CSR packing and the NumPy AP tail dominate"}}
    if with_cpu:
        from oracle.backend import OracleBackend          # checker / baseline only
        ob = OracleBackend()
        t0 = time.perf_counter()
        ob.pck(pa, **pkw)
        ob.map_assign(ma, **mkw)
        cpu_s = time.perf_counter() - t0
        res["cpu_port"] = {"value": n_frames / cpu_s, "unit": "frames/s", "s": cpu_s, "cores": 1,
                           "what": "C oracle (oracle/popnet_oracle.c) on the same CSR arrays"}
    return res


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from popnet_b200 import _abi, _lib, network, pipeline, synth
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.get()
    B = args.batch
    R = args.rotate
    model = network.rtpose_light3d(15, 14, 2, input_dim=1)
    sd = network.synth_state_dict(seed=0, style="reference")        # random-init weights of the architecture
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model.operand_dtype = _abi.OPERAND_BF16 if args.dtype == "bf16" else _abi.OPERAND_FP16
    est = pipeline.PoseEstimator(model, max_persons=32)
    # R rotating input sets so that consecutive steps never find their inputs in the 126 MB L2
    base = synth.depth_frames(min(B, 16), seed=1234 + 1000 * rank)
    heat, paf, depth, _ = synth.map_batch(B, seed=1234 + 1000 * rank, persons=(1, 6), noise=0.01)
    host_frames, dev_frames, dev_maps = [], [], []
    for r in range(R):
        fr = np.roll(np.tile(base, ((B + len(base) - 1) // len(base), 1, 1, 1))[:B], r, axis=0).copy()
        fr += np.float32(1e-3 * r)
        hf = torch.from_numpy(fr).pin_memory()
        host_frames.append(hf)
        dev_frames.append(hf.cuda())
        roll = lambda a: torch.from_numpy(np.roll(a, r, axis=0).copy()).cuda()
        dev_maps.append((roll(heat), roll(paf), roll(depth)))
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device(i, evs=None):
        """One step with inputs resident in HBM.  Forward on the main stream, decode + lift (+ all-gather) on the
        estimator's decode stream: consecutive steps are software-pipelined (decode of step i under the forward of
        step i+1), every step still runs the full forward and the full decode."""
        est.inject = dev_maps[i % R]
        x = dev_frames[i % R]
        est._buffers(B)
        slot = est._slots[i % est.NSLOT]

        def after(o):
            if evs is not None:
                evs[3].record(torch.cuda.current_stream())
            if world > 1:
                pipeline.gather_records(o, unpack=False)  # one NCCL all-gather of the packed record bytes
        out = est.infer_device(x, slot["out"], after=after, _evs=evs)
        return out

    def submit_e2e(i):
        est.inject = dev_maps[i % R]
        # H2D (copy stream) -> forward -> decode -> D2H (-> all-gather of the record bytes), asynchronous
        gather = (lambda o: pipeline.gather_records(o, unpack=False)) if world > 1 else None
        return est.submit(host_frames[i % R], after=gather)

    def run_e2e(n):
        """n steps through the public API, software-pipelined NSLOT deep: the H2D copy of step i+2 and the forward of step
        i+1 overlap the decode + D2H of step i; every step's records are read back on the host."""
        rec, q = None, collections.deque()
        for i in range(n):
            if len(q) == est.NSLOT:
                rec = est.collect(q.popleft())
            q.append(submit_e2e(i))
        while q:
            rec = est.collect(q.popleft())
        return rec

    # ---- value leg
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.popnet_launch_count()
    stage_evs = [[ev(), ev(), ev(), ev()] for _ in range(args.steps)]
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(args.steps):
        step_device(i, stage_evs[i])
    torch.cuda.current_stream().wait_stream(est.decode_stream)     # the timed region ends when the last decode has
    t1.record()
    barrier()
    launches = lib.popnet_launch_count() - launches0
    elapsed_ms = t0.elapsed_time(t1)
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in stage_evs]))     # forward, on its (main) stream
    dec_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in stage_evs]))     # decode + lift, on the decode stream
    # ---- e2e leg
    run_e2e(max(3, args.warmup // 2))
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
    # sanity: the timed path produced poses
    rec = run_e2e(1)
    n_person = int(np.asarray(rec["n_person"].cpu() if hasattr(rec["n_person"], "cpu") else rec["n_person"]).sum())
    flags = int(np.asarray(rec["flags"].cpu() if hasattr(rec["flags"], "cpu") else rec["flags"]).astype(np.int64).sum())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r1_forward_dram_traffic.json")
    if os.path.exists(tp) and B == 64:
        tj = json.load(open(tp))
        traffic, traffic_src = tj["dram_bytes_total"], "profiles/r1_forward_dram_traffic.json (ncu, one forward at batch 64)"
    frames_per_step = B * world
    value = frames_per_step * args.steps / (elapsed_ms * 1e-3)
    e2e = frames_per_step * args.steps / (e2e_ms * 1e-3)
    tflops = B * FLOP_PER_FRAME / (fwd_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": frames_per_step, "input": "224x224x1 fp32",
                   "weights": "random init of the architecture (reference init, seed 0)",
                   "decode_input": "GT-style maps (1-6 persons/frame, reference renderers' formulas) resident in HBM are "
                                   "decoded in place of the forward's own maps: untrained weights give sigma~0.5 heat-maps "
                                   "(thousands of plateau peaks), a degenerate decode workload (SURVEY.md 6.2); the forward "
                                   "still computes and writes all six maps",
                   "l2": "rotation of %d input sets (%.0f MB frames + %.0f MB maps) > 126 MB L2; activations (%.0f MB) "
                         "exceed L2 by themselves" % (R, R * B * 224 * 224 * 4 / 1e6, R * B * 46256 * 4 / 1e6,
                                                      lib.popnet_workspace_bytes(est.model._net_config(224, 224), B) / 1e6),
                   "parallelism": "dp%d (batch-sharded, all-gather of pose records)" % world},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": est.h2d_bytes(B), "d2h_bytes_per_step": est.d2h_bytes(B),
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (35 launches) + stem_kernel = the forward",
                     "achieved": tflops, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": tflops / pk["tflops"],
                     "frac_of_burst_peak": tflops / pk["tflops_burst"], "peak_source": pk["src"], "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_flop_per_step": B * FLOP_PER_FRAME,
                     "min_hbm_bytes_per_step": B * (224 * 224 * 4 + 2 * 46256 * 4) + 11_051_628,
                     "forward_ms": fwd_ms, "decode_ms": dec_ms,
                     "decode_hbm": {"bound": "hbm", "achieved": B * DECODE_BYTES_PER_FRAME / (dec_ms * 1e-3) / 1e9,
                                    "peak": pk["hbm"], "unit": "GB/s",
                                    "frac": B * DECODE_BYTES_PER_FRAME / (dec_ms * 1e-3) / 1e9 / pk["hbm"]}},
        "check": {"persons_decoded_per_step": n_person, "overflow_flags": flags},
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        fps, ts = cpu_path(64, 12, 1, cores)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "64 frames x 12 repeats: fp32 torch forward (%d threads) + C-oracle decode/lift "
                                          "over %d threads (%.1f s)" % (cores, cores, sum(ts))}
    if world == 1 and not args.no_evaluator:
        line["evaluator"] = evaluator_leg(with_cpu=not args.no_cpu_baseline)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--rotate", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-evaluator", action="store_true", help="skip the secondary evaluator metric")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
