run() { echo "$@"; env "$@" python tools/time_forward.py --batch 64 --iters 100 2>&1 | cut -c1-75 | tail -1; }
run A=0
run POPNET_SPLIT=80,38,30
run POPNET_SPLIT=74,37,37
run POPNET_SPLIT=84,36,28
run POPNET_SPLIT=71,53,24
run POPNET_SPLIT=106,42,0
run POPNET_SPLIT=72,40,36 POPNET_STAGE_NACC=2
run POPNET_SPLIT=72,40,36 POPNET_STAGE_NACC=3
run A=0
POPNET_SPLIT=80,38,30 python tools/forward_timeline.py 2>&1 | tail -32
