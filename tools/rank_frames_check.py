"""1-GPU diagnostic: for the bench frames of ranks 0..7, the pipelined (graph) estimator and an eager one must agree."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from popnet_b200 import network, pipeline  # noqa: E402

B = 64
model = network.rtpose_light3d(15, 14, 2, input_dim=1)
model.load_state_dict({k: torch.from_numpy(v) for k, v in bench.fixture_state_dict().items()})
est = pipeline.PoseEstimator(model, max_persons=32, strict=False)
solo = pipeline.PoseEstimator(model, max_persons=32, use_graphs=False, strict=False)
for r in range(8):
    fr = bench.make_frames(r, B, (1, 6))
    outs = []
    for e in (est, solo, est, solo):
        outs.append({k: np.array(v) for k, v in e.infer(fr).items()})
    ref = outs[1]
    n = ref["n_person"]
    for name, o in (("graph", outs[0]), ("graph again", outs[2]), ("eager again", outs[3])):
        bad = []
        for k in ref:
            for f in range(B):
                m = int(n[f])
                x, y = (o[k][f], ref[k][f]) if k in ("n_person", "flags") else (o[k][f, :m], ref[k][f, :m])
                if not np.array_equal(x, y, equal_nan=True):
                    bad.append((k, f))
        print("rank-%d frames: %-12s vs eager: %d mismatches %s  (persons %d, flags %d)" % (r, name, len(bad), bad[:4], int(n.sum()), int((ref["flags"] != 0).sum())), flush=True)
