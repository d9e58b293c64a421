#!/bin/bash
# usage: tools/sweep_wgrad.sh  -> prints ms per shape for several schedule settings (GPU box)
run() { echo "==== $*"; env "$@" python tools/bench_conv.py --kinds wgrad 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(f\"{r['shape']:24s} {r['ms']:8.4f} ms {r['tflops']:7.1f} TF\")
"; }
run A=1
run TPGAN_WGRAD_KB_MB=32
run TPGAN_WGRAD_KB_MB=128
run TPGAN_WGRAD_KB_MB_SLAB=64
run TPGAN_WGRAD_KB_MB_SLAB=100000
run TPGAN_WGRAD_SLAB_COLS=256
run TPGAN_WGRAD_SLAB_COLS=256 TPGAN_WGRAD_KB_MB_SLAB=64
