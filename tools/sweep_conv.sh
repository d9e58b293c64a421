#!/bin/bash
# usage: tools/sweep_conv.sh <dtype> <kinds> <only> VAR=VAL ...  -> one line per shape/kind (GPU box)
dt=$1; kinds=$2; only=$3; shift 3
echo "==== $dt $kinds $only $*"
env "$@" python tools/bench_conv.py --dtype $dt --kinds $kinds --only "$only" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(f\"{r['shape']:24s} {r['kind']:6s} {r['ms']:8.4f} ms {r['tflops']:7.1f} TF {r['frac_peak']:.3f}\")
"
