"""Multi-GPU diagnostic (torchrun): every rank runs one pipelined step on its own frames with the peer push; rank 0 compares
each rank's chunk of its gather buffer, field by field, with its own single-GPU decode of that rank's frames."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from popnet_b200 import _abi, network, p2p, pipeline  # noqa: E402
from popnet_b200._cuda_backend import records_layout  # noqa: E402
from popnet_b200.topology import MP3DHP, DecodeConfig  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
B = int(os.environ.get("B", 64))
calib = int(os.environ.get("CALIB", 0))
model = network.rtpose_light3d(15, 14, 2, input_dim=1)
model.load_state_dict({k: torch.from_numpy(v) for k, v in bench.fixture_state_dict().items()})
params = _abi.make_decode_params(DecodeConfig(), MP3DHP, max_persons=32, depth_channels=15)
peers = p2p.PeerGather(records_layout(B, params)[0], pipeline.PoseEstimator.NSLOT)
est = pipeline.PoseEstimator(model, max_persons=32, peers=peers, strict=False)
frames = torch.from_numpy(bench.make_frames(rank, B, (1, 6))).pin_memory()
if calib:
    print(rank, "calibration", est.calibrate(frames.cuda()), flush=True)
for rep in range(3):
    for _ in range(4):
        t = est.submit(frames)
        rec = est.collect(t)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in est.gathered(t).items()}
    mine = {k: np.array(v) for k, v in rec.items()}
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        solo = pipeline.PoseEstimator(model, max_persons=32, use_graphs=False, strict=False)
        bad = 0
        for r in range(world):
            ref = {k: np.array(v) for k, v in solo.infer(bench.make_frames(r, B, (1, 6))).items()}
            n = ref["n_person"]
            for k in ref:
                a = g[k][r * B:(r + 1) * B]
                for f in range(B):
                    m = int(n[f]) if k not in ("n_person", "flags") else None
                    x, y = (a[f], ref[k][f]) if m is None else (a[f, :m], ref[k][f, :m])
                    if not np.array_equal(x, y, equal_nan=True):
                        bad += 1
                        if bad <= 6:
                            print("rep %d: rank %d field %s frame %d: gathered %s vs solo %s" % (rep, r, k, f, np.asarray(x).ravel()[:6], np.asarray(y).ravel()[:6]), flush=True)
            if r == 0:
                same = all(np.array_equal(mine[k], ref[k], equal_nan=True) for k in ("n_person", "flags"))
                print("rep %d: rank 0 local records n_person/flags == solo: %s" % (rep, same), flush=True)
        print("rep %d: %d mismatching (rank, field, frame) entries" % (rep, bad), flush=True)
    dist.barrier()
peers.close()
dist.destroy_process_group()
