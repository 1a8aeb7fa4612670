"""Shared test helpers: golden fixture access and record comparison."""
import hashlib
import os

import numpy as np

from popnet_b200 import _abi, synth
from popnet_b200.topology import ITOP, MP3DHP, DecodeConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAMS = {"MP3DHP": MP3DHP, "ITOP": ITOP}
# must mirror tests/golden/make_golden.py::DECODE_CASES
DECODE_CASES = [("mp", 40, 100, (1, 6), 0.01, "MP3DHP"), ("crowd", 8, 900, (12, 16), 0.01, "MP3DHP"),
                ("itop", 6, 300, (1, 1), 0.02, "ITOP"), ("empty", 2, 500, (0, 0), 0.0, "MP3DHP")]


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return _cache[name]


def decode_case_inputs(case):
    name, B, seed, persons, noise, cam = case
    heat, paf, depth, _ = synth.map_batch(B, seed=seed, persons=persons, noise=noise)
    g = golden("decode_golden")
    if sha(heat, paf, depth) != str(g["%s/sha" % name]):
        # generator drifted (different NumPy?): fall back to the frames stored verbatim in the fixture
        heat, paf, depth = g["%s/maps_heat" % name], g["%s/maps_paf" % name], g["%s/maps_depth" % name]
    return heat, paf, depth


def params_for(camname, **kw):
    return _abi.make_decode_params(DecodeConfig(), CAMS[camname], **kw)


def compare_to_golden(out, f, g, key, K=15, exact=True, tol=2e-6):
    """Compare frame f of flat decode records with the reference's outputs stored under `key`.
    exact=True: every number bit-identical.  exact=False: ids/coordinates identical, scores within tol."""
    from popnet_b200.decode import records_to_reference
    jl, assoc = records_to_reference(out, f, K)
    gjl, gassoc = g[key + "joint_list"], g[key + "assoc"]
    assoc = np.asarray(assoc, np.float64).reshape(-1, K + 2)
    if jl.shape != gjl.shape or assoc.shape != gassoc.shape:
        return False, "shape %s/%s vs %s/%s" % (jl.shape, assoc.shape, gjl.shape, gassoc.shape)
    n = assoc.shape[0]
    p2, p3, pc = out["pose2d"][f, :n, :K], out["pose3d"][f, :n, :K], out["pose_conf"][f, :n, :K]
    if exact:
        ok = (np.array_equal(jl, gjl) and np.array_equal(assoc, gassoc) and np.array_equal(p2, g[key + "humans_2d"])
              and np.array_equal(p3, g[key + "humans_3d"]) and np.array_equal(pc, g[key + "conf"]))
        return ok, "bitwise mismatch"
    ok = (np.array_equal(jl[:, [0, 1, 3, 4]], gjl[:, [0, 1, 3, 4]]) and np.array_equal(assoc[:, :K], gassoc[:, :K])
          and np.array_equal(assoc[:, K + 1], gassoc[:, K + 1]) and np.array_equal(p2, g[key + "humans_2d"]))
    if ok:
        ok = (np.allclose(jl[:, 2], gjl[:, 2], atol=tol, rtol=0) and np.allclose(assoc[:, K], gassoc[:, K], atol=1e-4, rtol=0)
              and np.allclose(p3, g[key + "humans_3d"], atol=1e-5, rtol=0))
    return ok, "joint/assoc mismatch"


def records_equal(a, b, keys=None):
    """Byte-for-byte comparison of two flat record dicts on their VALID regions."""
    bad = []
    B = len(a["n_person"])
    for k in ("peak_count", "conn_count", "n_person", "flags"):
        if not np.array_equal(np.asarray(a[k]).astype(np.int64), np.asarray(b[k]).astype(np.int64)):
            bad.append(k)
    if bad:
        return bad
    for f in range(B):
        for t in range(a["peak_count"].shape[1]):
            n = a["peak_count"][f, t]
            for k in ("peak_xy", "peak_score"):
                if not np.array_equal(a[k][f, t, :n], b[k][f, t, :n]):
                    bad.append("%s[%d,%d]" % (k, f, t))
        for l in range(a["conn_count"].shape[1]):
            n = a["conn_count"][f, l]
            for k in ("conn_ij", "conn_score"):
                if not np.array_equal(a[k][f, l, :n], b[k][f, l, :n]):
                    bad.append("%s[%d,%d]" % (k, f, l))
        n = a["n_person"][f]
        for k in ("person_peak", "person_score", "person_njoint", "pose2d", "pose3d", "pose_conf"):
            if not np.array_equal(a[k][f, :n], b[k][f, :n]):
                bad.append("%s[%d]" % (k, f))
    return bad


def fixture_state_dict():
    """The fixture checkpoint (tests/golden/fixture_ckpt.npz): the reference module trained for a few thousand steps
    with the reference's own loss on synthetic frames (tools/make_fixture_ckpt.py).  Stored fp16, widened to fp32 here:
    the fixture IS the fp16-rounded weights, used identically by the reference, the oracle and the CUDA path."""
    z = np.load(os.path.join(GOLDEN, "fixture_ckpt.npz"))
    return {k: (z[k].astype(np.float32) if z[k].dtype != np.int64 else z[k]) for k in z.files}
