"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, the all-gather of pose records and
the all-reduce of evaluator counters must reproduce the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    from oracle.backend import OracleBackend
    from popnet_b200 import evaluate, pipeline, synth
    B = 12
    heat, paf, depth, _ = synth.map_batch(B, seed=321, persons=(1, 5), noise=0.01)
    params = helpers.params_for("MP3DHP", max_persons=32)
    sl = pipeline.shard(B, rank, world)
    local = c_oracle.decode(heat[sl], paf[sl], depth[sl], params)            # this rank's frames only
    from popnet_b200._cuda_backend import alloc_decode_out
    out = alloc_decode_out(sl.stop - sl.start, params, device="cpu")         # same packed layout as on the GPU
    for k in pipeline.RECORD_KEYS:
        src = local[k].view(np.int32) if local[k].dtype == np.uint32 else local[k]
        out[k].copy_(torch.from_numpy(np.ascontiguousarray(src)))
    full = pipeline.gather_records(out)
    # evaluator counters: shard the frames, all-reduce the integer counters
    ds = synth.eval_set(60, seed=5)
    evaluate._backend = OracleBackend()
    es = pipeline.shard(60, rank, world)
    part = evaluate.match_counts(ds["pred2d"][es], ds["gt2d"][es], pred3d=ds["pred3d"][es], gt3d=ds["gt3d"][es],
                                 num_joints=15, dist_th=0.1)
    red = pipeline.reduce_counts({"hit_cnt": part["hit_cnt"], "valid_cnt": part["valid_cnt"],
                                  "samples": np.array([part["samples_cnt"]])})
    # AP over frame shards of UNEQUAL size (60 frames, shards of 37 / 23): variable-length gather of the (score, label) rows
    names = ["j%d" % i for i in range(15)]
    cut = [0, 37, 60]
    ap_slice = slice(cut[rank], cut[rank + 1])
    evaluate.AP_TAIL = "numpy"
    ap = evaluate.eval_ap_3D_sharded(ds["pred3d"][ap_slice], ds["conf"][ap_slice], ds["gt3d"][ap_slice], [], names, thresh=0.1)
    if rank == 0:
        q.put(({k: np.asarray(v) for k, v in full.items()}, {k: v.numpy() for k, v in red.items()}, np.asarray(ap)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_and_reduce_match_single_process(oracle_lib):
    from oracle.backend import OracleBackend
    from popnet_b200 import evaluate, pipeline, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, red, ap = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    heat, paf, depth, _ = synth.map_batch(12, seed=321, persons=(1, 5), noise=0.01)
    single = oracle_lib.decode(heat, paf, depth, helpers.params_for("MP3DHP", max_persons=32))
    for k in pipeline.RECORD_KEYS:
        a = single[k].view(np.int32) if single[k].dtype == np.uint32 else single[k]
        assert np.array_equal(full[k], a), k
    ds = synth.eval_set(60, seed=5)
    evaluate._backend = OracleBackend()
    try:
        whole = evaluate.match_counts(ds["pred2d"], ds["gt2d"], pred3d=ds["pred3d"], gt3d=ds["gt3d"], num_joints=15, dist_th=0.1)
        evaluate.AP_TAIL = "numpy"
        ap_whole = evaluate.eval_ap_3D(ds["pred3d"], ds["conf"], ds["gt3d"], [], ["j%d" % i for i in range(15)], thresh=0.1)
    finally:
        evaluate._backend = None
        evaluate.AP_TAIL = "device"
    assert np.array_equal(ap, np.asarray(ap_whole)) and ap_whole[-1] > 0          # sharded AP == single-process AP, bit for bit
    assert np.array_equal(red["hit_cnt"], whole["hit_cnt"]) and np.array_equal(red["valid_cnt"], whole["valid_cnt"])
    assert int(red["samples"][0]) == whole["samples_cnt"]


def test_shard_covers_everything():
    from popnet_b200 import pipeline
    for n, w in ((512, 8), (10, 4), (3, 8), (0, 2)):
        idx = []
        for r in range(w):
            s = pipeline.shard(n, r, w)
            idx += list(range(n))[s]
        assert idx == list(range(n))
