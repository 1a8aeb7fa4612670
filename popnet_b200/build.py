"""Build the CUDA library IN-TREE (popnet_b200/libpopnet_b200.so) with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repository snapshot.  `python -m popnet_b200.build [--force] [--verbose]`.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpopnet_b200.so")
STAMP = os.path.join(HERE, "build", "stamp.txt")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr"]

# translation unit -> extra flags.  The decode and evaluator kernels reproduce NumPy / OpenCV rounding,
# so fused multiply-add contraction is off there (explicit fma() marks the BLAS-backed spots).
UNITS = {
    "abi.cu": [],
    "eval_kernels.cu": ["-fmad=false"],
    "decode_kernels.cu": ["-fmad=false"],
    "conv_kernels.cu": [],
    "forward.cu": [],
}


def _sources():
    return [u for u in UNITS if os.path.exists(os.path.join(CSRC, u))]


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")) and name != "packlists.c":
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode() + b"\0" + f.read())
    h.update(repr(sorted(UNITS.items())).encode())
    return h.hexdigest()


PACK_SRC = os.path.join(CSRC, "packlists.c")
PACK_OUT = os.path.join(HERE, "_packlists.so")


def build_packlists(force=False):
    """gcc -> popnet_b200/_packlists.so: the CPython extension that packs the evaluator's ragged lists (host code)."""
    import sysconfig
    if not force and os.path.exists(PACK_OUT) and os.path.getmtime(PACK_OUT) >= os.path.getmtime(PACK_SRC):
        return PACK_OUT
    cc = os.environ.get("CC", "gcc")
    cmd = [cc, "-O2", "-fPIC", "-shared", "-std=c11", "-Wall", "-I", sysconfig.get_paths()["include"], PACK_SRC, "-o", PACK_OUT]
    subprocess.run(cmd, check=True)
    return PACK_OUT


def build(force=False, verbose=False):
    build_packlists(force)
    digest = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return OUT
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = []
    procs = []
    for unit in _sources():
        obj = os.path.join(HERE, "build", unit.replace(".cu", ".o"))
        cmd = [NVCC, *ARCH, *COMMON, *UNITS[unit], "-c", os.path.join(CSRC, unit), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for unit, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            print("---- %s ----\n%s" % (unit, out), flush=True)
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [NVCC, *ARCH, "-shared", "-Xlinker", "--no-undefined", "-o", OUT, *objs, "-lcudart", "-lcuda"]
    subprocess.run(link, check=True)
    with open(STAMP, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
