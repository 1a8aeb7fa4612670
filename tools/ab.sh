run() { echo "$@"; env "$@" python tools/time_forward.py --batch 64 --iters 150 2>&1 | cut -c1-75 | tail -1; }
run POPNET_STEM_CPS=6
run POPNET_STEM_CPS=8
run POPNET_STEM_CPS=6
run POPNET_STEM_CPS=8
