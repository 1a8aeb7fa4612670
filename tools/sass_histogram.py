"""SASS opcode histogram of the built library (profiles/r2_sass_opcodes.txt): proves which tensor-core / bulk-copy
instructions the shipped kernels contain.  `python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt` (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "popnet_b200", "libpopnet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
ops, per_fn, fn = collections.Counter(), {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per_fn[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        ops[m.group(1)] += 1
        per_fn[fn][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCATOMSWS", "SYNCS", "HMMA", "ELECT", "UTCCP", "ACQBULK")
tot = collections.Counter()
for op, n in ops.items():
    tot[op.split(".")[0]] += n
print("# SASS opcode histogram of popnet_b200/libpopnet_b200.so (sm_100a): cuobjdump -sass, tools/sass_histogram.py")
print("# tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, tcgen05.alloc -> UTCATOMSWS, cp.async.bulk -> UBLKCP,")
print("# mbarrier -> SYNCS, elect.sync -> ELECT")
print("total instructions: %d in %d kernels\n" % (sum(ops.values()), len(per_fn)))
print("## tensor-core / bulk-copy / barrier opcodes (full mnemonics)")
for op, n in sorted(ops.items(), key=lambda t: -t[1]):
    if any(op.startswith(k) for k in KEY):
        print("%8d  %s" % (n, op))
print("\n## absent (library / previous-generation paths): HMMA=%d  UTMALDG=%d  (no mma.sync; no tensor-map TMA -- tiles are"
      " fetched with 1-D bulk copies, DESIGN.md section 3)\n" % (tot["HMMA"], tot["UTMALDG"]))
print("## per kernel")
names = subprocess.run(["c++filt"], input="\n".join(per_fn), capture_output=True, text=True).stdout.splitlines()
for fn, name in sorted(zip(per_fn, names), key=lambda t: t[1]):
    c = per_fn[fn]
    t = collections.Counter()
    for op, n in c.items():
        t[op.split(".")[0]] += n
    name = name.replace("(anonymous namespace)::", "").replace("popnet::", "").replace("void ", "")
    name = re.sub(r"\((?!anonymous).*", "", name)
    print("%-58s UTCHMMA %4d  UTCBAR %3d  LDTM %3d  UBLKCP %3d  instrs %5d" % (name[:58], t["UTCHMMA"], t["UTCBAR"], t["LDTM"], t["UBLKCP"], sum(c.values())))
print("\n## all opcode roots")
for root, n in sorted(tot.items(), key=lambda t: -t[1]):
    print("%8d  %s" % (n, root))
# PTX-level evidence (what the source asks for): count the tcgen05 / cp.async.bulk instructions in the .cu files
print("\n## source (inline PTX) mnemonics in popnet_b200/csrc/*.cu")
cnt = collections.Counter()
for f in sorted(os.listdir(os.path.join(ROOT, "popnet_b200", "csrc"))):
    if f.endswith((".cu", ".cuh")):
        for m in re.finditer(r"(tcgen05\.[a-z0-9_.:]+|cp\.async\.bulk[a-z0-9_.:]*|mbarrier\.[a-z0-9_.:]+|griddepcontrol\.[a-z_]+|elect\.sync)", open(os.path.join(ROOT, "popnet_b200", "csrc", f)).read()):
            cnt[m.group(1).rstrip(".:")] += 1
for k, n in sorted(cnt.items()):
    print("%4d  %s" % (n, k))
