#!/usr/bin/env python
"""Headline benchmark: depth frames/sec for rtpose_light3d forward + PAF decode + 3D lift (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c5] [--batch B | --global-batch G] [--dtype bf16|fp16]

N > 1 is launched by the driver through torchrun (one rank per GPU); rank 0 prints ONE JSON line.
A "step" is one pass of the hot path over one batch of synthetic depth frames per GPU:
    forward (39 convolutions on the tensor cores) -> decode of the network's OWN maps (peaks, limbs, assembly) -> 3D lift
    [-> pose records pushed into every rank's gather buffer over NVLink when N > 1].
Weights: the fixture checkpoint (tests/golden/fixture_ckpt.npz: the reference module trained with the reference's loss
on synthetic frames, tools/make_fixture_ckpt.py) -- no trained weights ship with the reference.
`value`   : inputs already resident in HBM, timed with CUDA events, max over ranks.
`e2e`     : the same step through the public API (popnet_b200.pipeline.PoseEstimator.submit / collect) with pinned HOST
            frames in and pose records copied back to the host every step.
The timed region is at least MIN_TIMED_S seconds whatever --steps is: the K steps are repeated `repeats` times inside it
(reported; ms_per_step is per step).
`--impl reference` times the CPU restatement of the reference's path (oracle/: fp32 torch forward + C decode of its maps)
on all host cores on the same workload -- the reference itself is Python and cannot travel to the GPU box.
Workloads: c2 = BASELINE.json configs[1] (batch 64 per GPU, 1-6 persons); c5 = configs[4] (crowded, 12-16 persons,
global batch 256); --global-batch 512 = configs[3] (C4: strong scaling, 512 frames sharded over the ranks).
"""
import argparse
import collections
import hashlib
import json
import math
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
MIN_TIMED_S = 2.0
WORKLOADS = {
    "c2": {"persons": (1, 6), "batch": 64, "max_persons": 32,
           "text": "C2: rtpose_light3d inference, batch 64 synthetic 224x224 depth frames per GPU (1-6 persons), "
                   "heatmaps+PAF+depth maps, greedy assembly + 3D lift of the network's own maps"},
    "c5": {"persons": (12, 16), "batch": 256, "max_persons": 64,
           "text": "C5: crowded scenes, 12-16 synthetic persons per frame with occlusion, global batch 256, "
                   "rtpose_light3d inference + decode + 3D lift of the network's own maps"},
}
FIXTURE = os.path.join(ROOT, "tests", "golden", "fixture_ckpt.npz")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", 1400.0), "tflops_burst": d.get("bf16_tflops", 1590.0),
                "hbm": d.get("hbm_gbs", 6650.0), "src": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


def fixture_state_dict():
    z = np.load(FIXTURE)
    return {k: (z[k].astype(np.float32) if z[k].dtype != np.int64 else z[k]) for k in z.files}


def make_frames(rank, B, persons, rot=0):
    """Rank `rank`'s batch of input set `rot`: up to 16 distinct seeded synthetic frames tiled to B (tile j offset by
    1e-3*j so no two frames are equal), rolled and offset per input set."""
    from popnet_b200 import synth
    base = synth.depth_frames(min(B, 16), seed=1234 + 1000 * rank, persons=persons)
    reps = (B + len(base) - 1) // len(base)
    fr = np.concatenate([base + np.float32(1e-3 * j) for j in range(reps)], 0)[:B]
    return (np.roll(fr, rot, axis=0) + np.float32(1e-3 * rot)).astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

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
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        ok = [r for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mhz = [float(r[0]) for r in ok]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in ok for n, v in zip(names, r[2:6]) if v == "Active"})
        mx = [float(r[1]) for r in ok if r[1].replace(".", "").isdigit()]
        pw = [float(r[6]) for r in ok if len(r) >= 7 and r[6].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(mhz), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_path(frames, repeats, warmup, cores, max_persons):
    """Returns (frames/s, seconds per repeat list).  forward: fp32 torch oracle with `cores` threads on the fixture
    checkpoint; decode + lift of ITS maps: C oracle, frames spread over `cores` threads (ctypes releases the GIL)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle, forward_torch
    from popnet_b200 import _abi
    from popnet_b200.topology import MP3DHP, DecodeConfig
    torch.set_num_threads(cores)
    sd = {k: torch.from_numpy(v) for k, v in fixture_state_dict().items()}
    x = torch.from_numpy(frames)
    n = len(frames)
    params = _abi.make_decode_params(DecodeConfig(), MP3DHP, max_persons=max_persons)
    c_oracle.lib()
    pool = ThreadPoolExecutor(cores)

    def one():
        (paf, heat, depth), _ = forward_torch.forward(sd, x)
        paf, heat, depth = paf.numpy(), heat.numpy(), depth.numpy()
        return list(pool.map(lambda f: c_oracle.decode(heat[f:f + 1], paf[f:f + 1], depth[f:f + 1], params), range(n)))

    for _ in range(warmup):
        one()
    ts = []
    for _ in range(repeats):
        t = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t)
    return n / float(np.mean(ts)), ts


def workload_config(args, world):
    wl = WORKLOADS[args.workload]
    if args.global_batch:
        B = args.global_batch // world
        scaling = "strong"
    elif args.workload == "c5":
        B = wl["batch"] // world
        scaling = "strong"
    else:
        B = args.batch or wl["batch"]
        scaling = "weak"
    return wl, max(B, 1), scaling


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl, B, scaling = workload_config(args, 1)
    # the same frames per step as the GPU arm's rank 0 unless that would run for more than a few minutes
    est_fps = 60.0
    budget_s = 150.0
    n = B
    while n > 4 and (args.steps + args.warmup) * n / est_fps > budget_s:
        n //= 2
    frames = make_frames(0, B, wl["persons"])[:n]
    fps, ts = cpu_path(frames, args.steps, args.warmup, cores, wl["max_persons"])
    desc = ("%d of the %d frames of a step, %d steps: fp32 torch forward (%d threads) + C-oracle decode/lift of its maps over "
            "%d threads; fixture checkpoint" % (n, B, args.steps, cores, cores))
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ts)) * 1e3 * B / n,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": bench_config(args, wl, B, max(1, args.gpus), scaling),      # the GPU arm's config at this N; the CPU sample is in cpu_baseline.sample
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bench_config(args, wl, B, world, scaling):
    return {"workload": wl["text"], "batch_per_gpu": B, "global_batch": B * world, "input": "224x224x1 fp32",
            "weights": "fixture checkpoint (reference module + reference loss on synthetic frames, tests/golden/fixture_ckpt.npz)",
            "decode_input": "the network's own output maps",
            "parallelism": "dp%d (batch-sharded; pose records pushed to every rank's gather buffer over NVLink)" % world}


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
        def __getattr__(self, name):
            return getattr(be, name)

        def pck(self, arrs, **kw):
            captured["pck"] = (arrs, kw)
            return be.pck(arrs, **kw)

        def map_assign(self, arrs, **kw):
            captured["map"] = (arrs, kw)
            return be.map_assign(arrs, **kw)

    def public_calls():
        with contextlib.redirect_stdout(io.StringIO()):
            evaluate.eval_human_dataset_3d(es["pred2d"], es["gt2d"], es["pred3d"], es["gt3d"], K, 0.1, 0.5)
            evaluate.eval_ap_3D(es["pred3d"], es["conf"], es["gt3d"], [], names, 0.1)

    evaluate._backend = Tap()
    public_calls()                                 # warm-up (library load, first-touch allocations)
    t0 = time.perf_counter()
    public_calls()
    e2e_s = time.perf_counter() - t0
    # the same two calls on Packed inputs -- what popnet_b200.io.load_results hands over when results and labels come
    # from files (main_evaluate_mp_human_3D.py's flow): the sets are packed once at load time, not once per call
    pk = {k: evaluate.Packed(*evaluate.pack_humans(es[k], K, 2 if k.endswith("2d") else 3)) for k in ("pred2d", "gt2d", "pred3d", "gt3d")}

    def packed_calls():
        with contextlib.redirect_stdout(io.StringIO()):
            evaluate.eval_human_dataset_3d(pk["pred2d"], pk["gt2d"], pk["pred3d"], pk["gt3d"], K, 0.1, 0.5)
            evaluate.eval_ap_3D(pk["pred3d"], es["conf"], pk["gt3d"], [], names, 0.1)

    packed_calls()
    t0 = time.perf_counter()
    packed_calls()
    packed_s = time.perf_counter() - t0
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
                   "note": "public reference-signature calls on ragged Python lists: list->CSR packing, H2D, kernels, AP tail, D2H"},
           "e2e_packed": {"value": n_frames / packed_s, "unit": "frames/s", "s": packed_s,
                          "note": "the same calls on Packed (CSR) inputs as popnet_b200.io.load_results produces them: H2D, kernels, AP tail, D2H"}}
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


def records_digest(rec, B_total):
    """sha256 over the VALID part of pose records (per frame: count, flags, then the rows of its persons)."""
    h = hashlib.sha256()
    n = np.asarray(rec["n_person"]).astype(np.int32)
    h.update(n.tobytes())
    h.update(np.asarray(rec["flags"]).astype(np.int32).tobytes())
    for f in range(B_total):
        m = int(n[f])
        for k in ("person_peak", "person_score", "person_njoint", "pose2d", "pose3d", "pose_conf"):
            h.update(np.ascontiguousarray(np.asarray(rec[k][f, :m])).tobytes())
    return h.hexdigest()


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from popnet_b200 import _abi, _lib, network, pipeline
    from popnet_b200._cuda_backend import records_layout
    from popnet_b200.topology import MP3DHP, DecodeConfig
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.get()
    wl, B, scaling = workload_config(args, world)
    model = network.rtpose_light3d(15, 14, 2, input_dim=1)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in fixture_state_dict().items()})
    model.operand_dtype = _abi.OPERAND_BF16 if args.dtype == "bf16" else _abi.OPERAND_FP16
    model.tuning = args.tuning
    peers = None
    if world > 1:
        from popnet_b200 import p2p
        params = _abi.make_decode_params(DecodeConfig(), MP3DHP, max_persons=wl["max_persons"], depth_channels=15)
        peers = p2p.PeerGather(records_layout(B, params)[0], pipeline.PoseEstimator.NSLOT)
    est = pipeline.PoseEstimator(model, max_persons=wl["max_persons"], peers=peers, use_graphs=not args.eager, strict=False)
    NS = est.NSLOT
    host_frames = [torch.from_numpy(make_frames(rank, B, wl["persons"], rot=r)).pin_memory() for r in range(args.rotate)]
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def agree(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t[0])

    # ---- SM split between the forward and the overlapped decode, chosen by the product's own calibration on the first
    # input set (part of the warm-up; crowded scenes need more decode SMs than the default 8)
    calib = {}
    if not args.no_calibrate:
        calib = est.calibrate(host_frames[0].cuda(non_blocking=True))
    # ---- value leg: the NSLOT slot input buffers hold NSLOT different input sets, resident in HBM
    for i in range(NS):
        est.slot_input(i, B).copy_(host_frames[i % len(host_frames)])
    torch.cuda.synchronize()

    def step_device(i, evs=None):
        return est.infer_device(est.slot_input(i, B), evs)

    launches0 = lib.popnet_launch_count()
    step_device(0)                                   # first use of slot 0: eager warm-up launches, then graph capture
    torch.cuda.synchronize()
    launches_per_step = int(lib.popnet_launch_count() - launches0) // (2 if est.use_graphs else 1)
    for i in range(1, max(args.warmup, NS)):
        step_device(i)
    barrier()
    # how long is a step?  (sets `repeats` so that the timed region is >= MIN_TIMED_S on every rank)
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(20):
        step_device(i)
    torch.cuda.current_stream().wait_stream(est.decode_stream)
    t1.record()
    torch.cuda.synchronize()
    est_ms = agree(t0.elapsed_time(t1) / 20, dist.ReduceOp.MIN if world > 1 else None)
    repeats = max(1, int(math.ceil(args.min_timed_s * 1e3 / (est_ms * args.steps))))
    total = args.steps * repeats
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(total):
        step_device(i)
    torch.cuda.current_stream().wait_stream(est.decode_stream)     # the timed region ends when the last decode has
    t1.record()
    barrier()
    elapsed_ms = t0.elapsed_time(t1)
    # forward / decode durations on their own streams (separate short loop: events per step cost host time)
    stage_evs = [[ev(), ev(), ev(), ev()] for _ in range(30)]
    for i in range(30):
        step_device(i, stage_evs[i])
    barrier()
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in stage_evs[5:]]))
    dec_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in stage_evs[5:]]))

    # ---- e2e leg: pinned host frames -> H2D -> forward -> decode -> D2H of the records, NSLOT batches in flight
    def run_e2e(n):
        rec, q = None, collections.deque()
        for i in range(n):
            if len(q) == NS:
                rec = est.collect(q.popleft())
            q.append(est.submit(host_frames[i % len(host_frames)]))
        while q:
            rec = est.collect(q.popleft())
        return rec

    run_e2e(max(NS, args.warmup))
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    run_e2e(total)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = agree(elapsed_ms, dist.ReduceOp.MAX if world > 1 else None)
    e2e_ms = agree(e2e_ms, dist.ReduceOp.MAX if world > 1 else None)

    # ---- checks: the timed path produced poses; multi-GPU: rank 0's gathered records == its own 1-GPU decode
    ticket = est.submit(host_frames[0])
    rec = est.collect(ticket)
    n_person = int(rec["n_person"].sum())
    flags = int((rec["flags"] != 0).sum())
    check = {"persons_decoded_per_step": n_person, "frames_with_overflow_flags": flags}
    if world > 1:
        torch.cuda.synchronize()
        gathered = {k: v.cpu().numpy() for k, v in est.gathered(ticket).items()}
        timed_out = peers.timed_out()
        barrier()
        if rank == 0:
            check["gather_sha"] = records_digest(gathered, B * world)
            solo = pipeline.PoseEstimator(model, max_persons=wl["max_persons"], use_graphs=False, strict=False)
            # (infer() returns views into one of the estimator's three slots: copy before the slot is reused)
            parts = [{k: np.array(v) for k, v in solo.infer(make_frames(r, B, wl["persons"], rot=0)).items()} for r in range(world)]
            cat = {k: np.concatenate([p[k] for p in parts], 0) for k in parts[0]}
            check["one_gpu_sha"] = records_digest(cat, B * world)
            check["gather_equals_one_gpu"] = check["gather_sha"] == check["one_gpu_sha"]
        check["p2p_timed_out"] = bool(timed_out)
        barrier()
        # evaluator over frame shards on NCCL: all-reduced PCK counters and the AP of the variable-length (score, label) gather
        # must equal rank 0's single-GPU evaluation of the whole set
        import contextlib
        import io as _io
        from popnet_b200 import evaluate, synth
        ds = synth.eval_set(400, seed=9)
        names = ["j%d" % i for i in range(15)]
        cut = [round(400 * r / world * (0.9 if 0 < r < world else 1.0)) for r in range(world + 1)]     # unequal shards
        sl = slice(cut[rank], cut[rank + 1])
        with contextlib.redirect_stdout(_io.StringIO()):
            part = evaluate.match_counts(ds["pred2d"][sl], ds["gt2d"][sl], pred3d=ds["pred3d"][sl], gt3d=ds["gt3d"][sl],
                                         num_joints=15, dist_th=0.1)
            red = pipeline.reduce_counts({"hit_cnt": torch.as_tensor(part["hit_cnt"]).cuda(),
                                          "valid_cnt": torch.as_tensor(part["valid_cnt"]).cuda()})
            ap = evaluate.eval_ap_3D_sharded(ds["pred3d"][sl], ds["conf"][sl], ds["gt3d"][sl], [], names, thresh=0.1)
            if rank == 0:
                whole = evaluate.match_counts(ds["pred2d"], ds["gt2d"], pred3d=ds["pred3d"], gt3d=ds["gt3d"], num_joints=15, dist_th=0.1)
                ap1 = evaluate.eval_ap_3D(ds["pred3d"], ds["conf"], ds["gt3d"], [], names, thresh=0.1)
        if rank == 0:
            check["evaluator_sharded"] = {
                "pck_counters_equal_one_gpu": bool(np.array_equal(red["hit_cnt"].cpu().numpy(), whole["hit_cnt"]) and
                                                   np.array_equal(red["valid_cnt"].cpu().numpy(), whole["valid_cnt"])),
                "ap_equals_one_gpu": bool(np.array_equal(np.asarray(ap), np.asarray(ap1))), "mAP": float(ap1[-1])}
        barrier()
    if rank != 0:
        if world > 1:
            peers.close()
            dist.destroy_process_group()
        return
    pk = peaks()
    traffic, traffic_src = None, None
    for name in ("r2_forward_dram_traffic.json", "r1_forward_dram_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp) and B == 64:
            tj = json.load(open(tp))
            traffic, traffic_src = tj["dram_bytes_total"], "profiles/%s (ncu, one forward at batch 64)" % name
            break
    frames_per_step = B * world
    value = frames_per_step * total / (elapsed_ms * 1e-3)
    e2e = frames_per_step * total / (e2e_ms * 1e-3)
    tflops = B * FLOP_PER_FRAME / (fwd_ms * 1e-3) / 1e12
    sustained = elapsed_ms >= 1000.0
    peak = pk["tflops"] if sustained else pk["tflops_burst"]
    cfg = bench_config(args, wl, B, world, scaling)
    cfg["l2"] = ("%d rotating input sets; one step moves > 1.5 GB through the 126 MB L2 (activation workspace %.0f MB), "
                 "nothing survives from one step to the next" % (NS, lib.popnet_workspace_bytes(model._net_config(224, 224), B) / 1e6))
    # launch schedule of this run (implementation detail, kept out of `config`, which names the workload only)
    schedule = {"cuda_graphs": bool(est.use_graphs), "decode_sms": int(est.reserve_sms)}
    if calib:
        schedule["calibration_ms_per_step"] = {str(k): round(v, 4) for k, v in calib.items()}
    if args.tuning:
        schedule["tuning"] = "0x%x" % args.tuning
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "repeats": repeats, "timed_region_s": elapsed_ms * 1e-3,
        "ms_per_step": elapsed_ms / total, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": cfg, "schedule": schedule,
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": est.h2d_bytes(B), "d2h_bytes_per_step": est.d2h_bytes(B),
                "ms_per_step": e2e_ms / total, "timed_region_s": e2e_ms * 1e-3},
        "gpu_launches": int(launches_per_step * total),
        "gpu_launches_per_step": launches_per_step,
        "host_launches_per_step": 2 if est.use_graphs else launches_per_step,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel + stem_kernel = the forward",
                     "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak,
                     "peak_regime": "sustained (timed region %.1f s)" % (elapsed_ms * 1e-3) if sustained else "burst",
                     "frac_of_sustained_peak": tflops / pk["tflops"], "frac_of_burst_peak": tflops / pk["tflops_burst"],
                     "peak_source": pk["src"], "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_flop_per_step": B * FLOP_PER_FRAME,
                     "min_hbm_bytes_per_step": B * (224 * 224 * 4 + 2 * 46256 * 4) + 11_051_628,
                     "forward_ms": fwd_ms, "decode_ms": dec_ms,
                     "decode_hbm": {"bound": "hbm", "achieved": B * DECODE_BYTES_PER_FRAME / (dec_ms * 1e-3) / 1e9,
                                    "peak": pk["hbm"], "unit": "GB/s",
                                    "frac": B * DECODE_BYTES_PER_FRAME / (dec_ms * 1e-3) / 1e9 / pk["hbm"]}},
        "check": check,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = min(B, 64)
        reps = 12 if args.workload == "c2" else 4
        fps, ts = cpu_path(make_frames(0, B, wl["persons"])[:n], reps, 1, cores, wl["max_persons"])
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "%d frames x %d repeats: fp32 torch forward (%d threads) + C-oracle decode/lift of its maps "
                                          "over %d threads (%.1f s); fixture checkpoint" % (n, reps, cores, cores, sum(ts))}
    if world == 1 and not args.no_evaluator:
        line["evaluator"] = evaluator_leg(with_cpu=not args.no_cpu_baseline)
    print(json.dumps(line), flush=True)
    if world > 1:
        peers.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (weak scaling; default: the workload's)")
    ap.add_argument("--global-batch", type=int, default=0, help="total frames per step, sharded over the ranks (strong scaling; C4: 512)")
    ap.add_argument("--dtype", default="fp16", choices=["bf16", "fp16"],
                    help="16-bit operand format (fp32 accumulate); fp16 is the product default, see DESIGN.md section 2")
    ap.add_argument("--rotate", type=int, default=3)
    ap.add_argument("--eager", action="store_true", help="issue every launch from the host instead of replaying CUDA graphs")
    ap.add_argument("--tuning", type=lambda v: int(v, 0), default=0, help="PopnetNetConfig.tuning bits (A/B of launch schedules; 0 = product defaults)")
    ap.add_argument("--min-timed-s", type=float, default=MIN_TIMED_S, help="lower bound of the timed regions (0 under a profiler)")
    ap.add_argument("--no-calibrate", action="store_true", help="keep the default 8 decode SMs instead of PoseEstimator.calibrate()")
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
