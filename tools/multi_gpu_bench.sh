#!/bin/bash
# The three multi-GPU bench lines (weak C2, C4 = 512 frames sharded, C5 = crowded, 256 frames sharded) on NG GPUs of one box.
mkdir -p gpurun_out
NG=${NG:-8}
run() {
  n=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $NG "$@" > gpurun_out/$n.json 2> gpurun_out/$n.err
  echo "$n rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/$n.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 4), d["config"].get("global_batch"), d["scaling"], d.get("schedule", {}).get("decode_sms"), d["check"], d["clocks"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$n.err").read()[-1500:])
PY
}
run bench_n$NG
run bench_c4_global512_n$NG --global-batch 512 --steps 50
run bench_c5_n$NG --workload c5 --steps 50
