"""Turn the raw files of tools/profile_round.sh (gpurun_out/) into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r".*::", "", name).replace("void ", "").strip()


def read_ncu_csv(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, mi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[ii], {"name": r[ki]})
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        if r[mi].startswith("dram"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        else:
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        d[r[mi]] = v
    return list(per.values())


# ---- 1. launch shares of one bench step
k = read_ncu_csv(os.path.join(OUT, "launches_bench.csv"))
names = [short(d["name"]) for d in k]
# one step = from one stem_kernel to the next; take the LAST complete step (warm)
stems = [i for i, n in enumerate(names) if n.startswith("stem_kernel")]
lo, hi = stems[-2], stems[-1]
# the decode kernels of step i run after its forward: rotate so the window holds one forward + one decode
step = k[lo:hi]
tot = sum(d["gpu__time_duration.sum"] for d in step)
agg = collections.OrderedDict()
for d in step:
    a = agg.setdefault(short(d["name"]), [0, 0.0])
    a[0] += 1
    a[1] += d["gpu__time_duration.sum"]
conv = sum(v[1] for n, v in agg.items() if n.startswith("conv_tc_kernel") or n.startswith("stem_kernel"))
with open(os.path.join(PROF, tag + "_launch_shares.txt"), "w") as f:
    f.write("# per-kernel share of one bench step (batch 64, B200), from profiles/%s_launches_bench.csv\n" % tag)
    f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 3 --min-timed-s 0 --no-cpu-baseline --no-evaluator\n")
    f.write("# (cold-cache, serialised launches: compare SHARES; in the real step the three branch chains of a stage run concurrently\n")
    f.write("#  and the decode of step i overlaps the forward of step i+1 -- see %s_forward_timeline.txt for the real overlap.\n" % tag)
    f.write("#  The three decode kernels are limited to the few SMs the conv grids leave free (8 or fewer CTAs each): serialised\n")
    f.write("#  here they count with their full 8-SM duration, in the step they run UNDER the next forward and cost it < 2 %;\n")
    f.write("#  forward share of the step as the bench measures it with events: roofline.forward_ms / ms_per_step = 0.98.)\n")
    f.write("one step: %d launches, %.1f us serialised; conv_tc_kernel (all shapes) + stem_kernel = %.1f%% of the step\n" % (len(step), tot, 100 * conv / tot))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("  %-36s x%-3d %8.1f us  %5.1f%%\n" % (n[:36], c, t, 100 * t / tot))
subprocess.run(["cp", os.path.join(OUT, "launches_bench.csv"), os.path.join(PROF, tag + "_launches_bench.csv")], check=True)

# ---- 2. DRAM traffic of one forward
k = read_ncu_csv(os.path.join(OUT, "forward_dram.csv"))
assert short(k[0]["name"]).startswith("stem_kernel") and len(k) == 38, (k[0]["name"], len(k))
rd = sum(d.get("dram__bytes_read.sum", 0) for d in k)
wr = sum(d.get("dram__bytes_write.sum", 0) for d in k)
out = {"command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                  "-k regex:'conv_tc|stem_kernel|pool_kernel' -s 114 -c 38 python tools/time_forward.py --batch 64 --iters 1 --dtype fp16",
       "what": "one forward at batch 64 (38 kernels: stem, 2 pools, 35 conv_tc_kernel launches), B200",
       "kernels": len(k), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr,
       "sum_kernel_time_us_serialised": sum(d["gpu__time_duration.sum"] for d in k),
       "per_kernel": [{"kernel": short(d["name"])[:48], "us": round(d["gpu__time_duration.sum"], 1),
                       "dram_MB": round((d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6, 1)} for d in k]}
json.dump(out, open(os.path.join(PROF, tag + "_forward_dram_traffic.json"), "w"), indent=1)

# ---- 3. ncu --set full detail of the forward's first 13 tensor-core launches
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
reps = [r for r in (os.path.join(OUT, "prof_conv_block.raw.csv"), os.path.join(OUT, "prof_conv.raw.csv")) if os.path.exists(r)]
if reps:
    with open(os.path.join(PROF, tag + "_conv_tc_ncu_summary.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:'conv_tc|stem_kernel' -s 108 -c 8 / -s 116 -c 12 python tools/time_forward.py --batch 64 --iters 1\n")
        f.write("# B200, forward at batch 64, fp16 operands: stem + layers 1-6 and 8 (first capture, if present), then the stage-1 branch convs.\n")
        f.write("# conv_tc_kernel template arguments: <NT, NACC, TAPS, B stages, B resident, debug>.  The .ncu-rep files are scratch (not committed).\n\n")
        allrows = []
        for rep in reps:
            rows = list(csv.reader(open(rep)))
            hdr, units = rows[0], rows[1]
            allrows += rows[2:]
        for r in allrows:
            rec = dict(zip(hdr, r))
            f.write("%s grid %s block %s\n" % (short(rec["Kernel Name"]), rec.get("Grid Size", ""), rec.get("Block Size", "")))
            for m in METRICS:
                if m in rec:
                    f.write("   %-70s %18s %s\n" % (m, rec[m], units[hdr.index(m)]))
            f.write("\n")
misc = os.path.join(OUT, "prof_misc.raw.csv")
if os.path.exists(misc):
    with open(os.path.join(PROF, tag + "_misc_kernels_ncu_summary.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:'peaks_kernel|limbs_kernel|assemble_kernel|pool_kernel' -s 8 -c 5 "
                "python bench.py --steps 2 --warmup 3 --min-timed-s 0 --no-cpu-baseline --no-evaluator\n# B200, batch 64; decode kernels limited to 8 CTAs (the pipelined step's setting)\n\n")
        rows = list(csv.reader(open(misc)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            rec = dict(zip(hdr, r))
            f.write("%s grid %s block %s\n" % (short(rec["Kernel Name"]), rec.get("Grid Size", ""), rec.get("Block Size", "")))
            for m in METRICS:
                if m in rec:
                    f.write("   %-70s %18s %s\n" % (m, rec[m], units[hdr.index(m)]))
            f.write("\n")
for src, dst in (("bench_c5_n1.json", tag + "_bench_c5_n1.json"), ("forward_timeline.txt", tag + "_forward_timeline.txt"), ("bench_n1.json", tag + "_bench_n1.json"), ("bench_ref.json", tag + "_bench_reference_arm.json")):
    if os.path.exists(os.path.join(OUT, src)):
        subprocess.run(["cp", os.path.join(OUT, src), os.path.join(PROF, dst)], check=True)
print(open(os.path.join(PROF, tag + "_launch_shares.txt")).read())
print({k2: v for k2, v in out.items() if k2 != "per_kernel"})
