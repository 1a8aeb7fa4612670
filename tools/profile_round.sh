#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): bench line, reference arm, launch list, DRAM traffic, timeline,
# ncu detail.  tools/summarize_profiles.py r2 turns the raw files into profiles/r2_*.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python bench.py --workload c5 --no-cpu-baseline --no-evaluator --steps 50 > gpurun_out/bench_c5_n1.json 2>/dev/null
python tools/forward_timeline.py > gpurun_out/forward_timeline.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --min-timed-s 0 --no-cpu-baseline --no-evaluator > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc|stem_kernel|pool_kernel" -s 114 -c 38 --csv --log-file gpurun_out/forward_dram.csv python tools/time_forward.py --batch 64 --iters 1 --dtype fp16 > /dev/null 2>&1
# 36 conv_tc/stem launches per forward: 108 = start of the 4th forward (stem, layers 1-6, 8); 116 = its stage-1 branches
ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem_kernel" -s 108 -c 8 -o gpurun_out/prof_conv_block python tools/time_forward.py --batch 64 --iters 1 --dtype fp16 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem_kernel" -s 116 -c 12 -o gpurun_out/prof_conv python tools/time_forward.py --batch 64 --iters 1 --dtype fp16 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"peaks_kernel|limbs_kernel|assemble_kernel|pool_kernel" -s 8 -c 5 -o gpurun_out/prof_misc python bench.py --steps 2 --warmup 3 --min-timed-s 0 --no-cpu-baseline --no-evaluator > /dev/null 2>&1
# the .ncu-rep files are ~40 MB each and gpurun only brings back 64 MiB: keep the raw-page CSVs instead
for r in prof_conv_block prof_conv prof_misc; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null && rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out
