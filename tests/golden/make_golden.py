"""Generate the committed golden fixtures by running the UNMODIFIED reference (imported read-only
from /root/reference through tests/golden/refshim.py) on seeded synthetic inputs.

Run in the build container only:   python tests/golden/make_golden.py [decode] [eval] [forward]
Outputs (committed):  tests/golden/decode_golden.npz, eval_golden.npz, forward_golden.npz
The reference ships no golden vectors or tests; these files are what pins the
oracle (oracle/) and, through it, the CUDA path.
"""
import contextlib
import copy
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE]

import refshim  # noqa: E402
from popnet_b200 import synth  # noqa: E402
from popnet_b200.topology import JOINT_NAMES, MP3DHP, ITOP  # noqa: E402

DECODE_CASES = [
    # name, batch, seed, persons, noise, camera
    ("mp", 40, 100, (1, 6), 0.01, "MP3DHP"),
    ("crowd", 8, 900, (12, 16), 0.01, "MP3DHP"),
    ("itop", 6, 300, (1, 1), 0.02, "ITOP"),
    ("empty", 2, 500, (0, 0), 0.0, "MP3DHP"),
]
CAMS = {"MP3DHP": MP3DHP, "ITOP": ITOP}


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def decode_inputs(case):
    name, B, seed, persons, noise, cam = case
    heat, paf, depth, _ = synth.map_batch(B, seed=seed, persons=persons, noise=noise)
    return heat, paf, depth


def make_decode():
    import cv2
    ref = refshim.load()
    out = {}
    for case in DECODE_CASES:
        name, B, seed, persons, noise, camname = case
        cam = CAMS[camname]
        heat, paf, depth = decode_inputs(case)
        out["%s/sha" % name] = np.array(sha(heat, paf, depth))
        # keep the first two frames' maps verbatim so the fixtures survive a change of the generator
        out["%s/maps_heat" % name] = heat[:2]
        out["%s/maps_paf" % name] = paf[:2]
        out["%s/maps_depth" % name] = depth[:2]
        for variant, ipp in (("native", False), ("ipp", True)):
            cv2.ipp.setUseIPP(ipp)
            for f in range(B):
                r = refshim.reference_decode_frame(ref, heat[f], paf[f], depth[f], cam)
                P = len(r["humans_2d"])
                key = "%s/%s/%d/" % (name, variant, f)
                out[key + "joint_list"] = r["joint_list"]
                out[key + "assoc"] = r["assoc"]
                out[key + "humans_2d"] = np.asarray(r["humans_2d"], np.float64).reshape(P, 15, 2)
                out[key + "humans_3d"] = np.asarray(r["humans_3d"], np.float64).reshape(P, 15, 3)
                out[key + "conf"] = np.asarray(r["conf"], np.float64).reshape(P, 15)
        cv2.ipp.setUseIPP(True)
        print("decode case", name, "frames", B)
    np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), **out)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def make_eval():
    import warnings
    warnings.simplefilter("ignore")
    ref = refshim.load()
    names = list(JOINT_NAMES)
    out = {}
    for tag, N, seed in (("small", 400, 7), ("c3", 4000, 0)):
        ds = synth.eval_set(N, seed=seed)
        flat = [np.asarray([h for fr in ds[k] for h in fr], np.float64) for k in ("pred2d", "pred3d", "conf", "gt2d", "gt3d")]
        out[tag + "/sha"] = np.array(sha(*flat))
        th2d = 0.02 * np.sqrt(480 ** 2 + 512 ** 2)
        a, k = quiet(ref.eval_pck.eval_human_dataset_2d_PCKh, ds["pred2d"], ds["gt2d"], 0, 1, 15, 0.5, 0.5)
        out[tag + "/pckh_avg"], out[tag + "/pckh_kcp"] = np.asarray(a), np.asarray(k)
        a, k = quiet(ref.eval_pck.eval_human_dataset_2d, ds["pred2d"], ds["gt2d"], 15, th2d, 0.5)
        out[tag + "/pck2d_avg"], out[tag + "/pck2d_kcp"] = np.asarray(a), np.asarray(k)
        a, k = quiet(ref.eval_pck.eval_human_dataset_3d, ds["pred2d"], ds["gt2d"], ds["pred3d"], ds["gt3d"], 15, 0.1, 0.5)
        out[tag + "/pck3d_avg"], out[tag + "/pck3d_kcp"] = np.asarray(a), np.asarray(k)
        out[tag + "/ap2d"] = quiet(ref.eval_mAP.eval_ap_mpii_v2, ds["pred2d"], copy.deepcopy(ds["conf"]), ds["gt2d"], [], 0, 1, names, 0.5)
        out[tag + "/ap3d"] = quiet(ref.eval_mAP.eval_ap_3D, ds["pred3d"], copy.deepcopy(ds["conf"]), ds["gt3d"], [], names, 0.1)
        # integer contract + per-GT distances, through the reference's own building blocks
        d2, d3 = [], []
        for f in range(N):
            if len(ds["gt2d"][f]) == 0:
                continue
            d2 += ref.eval_pck.match_humans_2d(ds["pred2d"][f], ds["gt2d"][f], 0.5)
            d3 += ref.eval_pck.match_humans_3d(ds["pred2d"][f], ds["gt2d"][f], ds["pred3d"][f], ds["gt3d"][f], 0.5)
        d2, d3 = np.asarray(d2), np.asarray(d3)
        hsz = np.asarray([h for fr in ds["gt2d"] for h in ref.eval_pck.compute_head_size(fr, 0, 1)])
        out[tag + "/hit_pckh"] = np.sum((d2 >= 0) & (d2 < hsz[:, None] * 0.5), 0).astype(np.int64)
        out[tag + "/hit_pck2d"] = np.sum((d2 >= 0) & (d2 < th2d), 0).astype(np.int64)
        out[tag + "/hit_pck3d"] = np.sum((d3 >= 0) & (d3 < 0.1), 0).astype(np.int64)
        out[tag + "/valid2d"] = np.sum(d2 >= 0, 0).astype(np.int64)
        out[tag + "/valid3d"] = np.sum(d3 >= 0, 0).astype(np.int64)
        out[tag + "/samples"] = np.array(d2.shape[0], np.int64)
        for dim, key, refd in ((2, "pred2d", None), (3, "pred3d", None)):
            gtk = "gt2d" if dim == 2 else "gt3d"
            vis = [np.ones((len(g), 15)).tolist() for g in ds[gtk]]
            if dim == 2:
                rd = [ref.eval_mAP.compute_head_size_from_two_joints(g, 0, 1) for g in ds[gtk]]
                th = 0.5
            else:
                rd = [np.ones(len(g)).tolist() for g in ds[gtk]]
                th = 0.1
            sc, lb, ngt = ref.eval_mAP.assignGTmulti(ds[key], ds["conf"], ds[gtk], vis, rd, 15, th)
            out[tag + "/map%d_npos" % dim] = np.array([sum(int(x) for fr in lb[j] for x in fr) for j in range(15)], np.int64)
            out[tag + "/map%d_nscores" % dim] = np.array([sum(len(fr) for fr in sc[j]) for j in range(15)], np.int64)
            out[tag + "/map%d_ngt" % dim] = ngt.sum(1).astype(np.int64)
            if tag == "small":
                out[tag + "/map%d_labels" % dim] = np.array([[int(x) for fr in lb[j] for x in fr] for j in range(15)], np.uint8).T
        if tag == "small":
            out[tag + "/dists2d"], out[tag + "/dists3d"] = d2, d3
        print("eval case", tag, "GT humans", d2.shape[0])
    np.savez_compressed(os.path.join(HERE, "eval_golden.npz"), **out)


def make_forward():
    """Reference module (its own forward, fp32, CPU) on the C1 frame and two synthetic frames, for two
    seeded checkpoints built by popnet_b200.network.synth_state_dict (no trained weights ship)."""
    import cv2
    import torch
    from popnet_b200 import network
    ref = refshim.load()
    out = {}
    # C1 input: bundled ITOP frame through the ITOP eval-time recipe
    # (datasets_itop_rtpose.py:213-223: resize 224 bilinear, clamp [0, 5], (x-3)/2)
    raw = np.load(os.path.join(refshim.REFERENCE_ROOT, "third_party_methods", "00_02254.npy")).astype(np.float32)
    img = cv2.resize(raw, (224, 224), interpolation=cv2.INTER_LINEAR)
    img = np.clip(img, 0, 5.0)
    x = np.concatenate([((img - 3.0) / 2.0).astype(np.float32)[None, None], synth.depth_frames(2, seed=4321)], 0)
    out["x"] = x.astype(np.float16)          # stored as fp16; the test feeds exactly these (rounded) frames
    x = out["x"].astype(np.float32)
    torch.set_num_threads(8)
    for style in ("reference", "trained_like"):
        sd = network.synth_state_dict(seed=11, style=style)
        out[style + "/state_sha"] = np.array(sha(*[sd[k] for k in sorted(sd)]))
        model = ref.rtpose_light3d(15, 14, 2, input_dim=1).float().eval()
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        with torch.no_grad():
            (paf, heat, depth), saved = model(torch.from_numpy(x))
        out[style + "/paf"], out[style + "/heat"], out[style + "/depth"] = paf.numpy(), heat.numpy(), depth.numpy()
        out[style + "/paf1"], out[style + "/heat1"], out[style + "/depth1"] = (t.numpy() for t in saved[:3])
        print("forward style", style)
    np.savez_compressed(os.path.join(HERE, "forward_golden.npz"), **out)


E2E_FRAMES, E2E_SEED, E2E_MAP_FRAMES = 1024, 777_000, 8


def make_e2e():
    """End-to-end golden of the fixture checkpoint (tests/golden/fixture_ckpt.npz, tools/make_fixture_ckpt.py): the
    reference's eval loop (evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:161-263) on 1024 seeded synthetic depth
    frames -- reference module forward (fp32, CPU) -> paf_to_pose -> paf_to_human_list -> retrieve_depth_heat_weighted ->
    rescale / back-projection.  Stored: per frame the assembled persons (2D joints, 3D joints, confidences) and the
    peak / person counts; the six fp32 output maps of the first 8 frames (tolerance anchor)."""
    import cv2
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    ref = refshim.load()
    sd = helpers.fixture_state_dict()
    model = ref.rtpose_light3d(15, 14, 2, input_dim=1).float().eval()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    torch.set_num_threads(os.cpu_count() or 1)
    x = synth.depth_frames(E2E_FRAMES, seed=E2E_SEED)
    out = {"x_sha": np.array(sha(x)), "state_sha": np.array(sha(*[sd[k] for k in sorted(sd)]))}
    n_person, off, p2, p3, pc, npk = [], [0], [], [], [], []
    cv2.ipp.setUseIPP(False)                 # OpenCV's own C++ resize: the path the oracle is bit-exact against
    for b0 in range(0, E2E_FRAMES, 32):
        with torch.no_grad():
            (paf, heat, depth), saved = model(torch.from_numpy(x[b0:b0 + 32]))
        paf, heat, depth = paf.numpy(), heat.numpy(), depth.numpy()
        if b0 == 0:
            for name, t in (("paf", paf), ("heat", heat), ("depth", depth)):
                out["maps/" + name] = t[:E2E_MAP_FRAMES].copy()
            for name, t in zip(("paf1", "heat1", "depth1"), saved[:3]):
                out["maps/" + name] = t.numpy()[:E2E_MAP_FRAMES].copy()
        for f in range(paf.shape[0]):
            r = refshim.reference_decode_frame(ref, heat[f], paf[f], depth[f], MP3DHP)
            P = len(r["humans_2d"])
            n_person.append(P)
            npk.append(len(r["joint_list"]))
            off.append(off[-1] + P)
            p2.append(np.asarray(r["humans_2d"], np.float64).reshape(P, 15, 2))
            p3.append(np.asarray(r["humans_3d"], np.float64).reshape(P, 15, 3))
            pc.append(np.asarray(r["conf"], np.float64).reshape(P, 15))
        print("e2e frames", b0 + 32, "persons so far", off[-1], flush=True)
    cv2.ipp.setUseIPP(True)
    out["n_person"] = np.asarray(n_person, np.int32)
    out["n_peaks"] = np.asarray(npk, np.int32)
    out["off"] = np.asarray(off, np.int32)
    out["pose2d"] = np.concatenate(p2, 0)
    out["pose3d"] = np.concatenate(p3, 0).astype(np.float32)     # compared with a tolerance (bf16 forward): fp32 storage
    out["conf"] = np.concatenate(pc, 0).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "e2e_golden.npz"), **out)
    print("e2e golden:", E2E_FRAMES, "frames,", off[-1], "persons")


if __name__ == "__main__":
    what = sys.argv[1:] or ["decode", "eval", "forward"]
    if "e2e" in what:
        make_e2e()
    if "decode" in what:
        make_decode()
    if "eval" in what:
        make_eval()
    if "forward" in what:
        make_forward()
