#!/bin/bash
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  echo "=== $label"
  for sh in enh64 enh32 enh16 add64 conv1_64 up128 up32 conv4rb; do env "$@" python tools/bench_conv.py --only $sh --kinds fwd,dgrad | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  %-22s %-5s %.4f ms %6.1f TF' % (d['shape'], d['kind'], d['ms'], d['tflops']))"; done
  env TPGAN_FLATCONV=0 "$@" python tools/bench_local.py --cin 128 --cout 128 --div 2 --kinds fwd | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  local128 fwd %.4f ms graph %.4f' % (d['ms'], d['graph_ms']))"
  env "$@" python bench.py --no-cpu --no-secondary 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  bench', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
}
run default X=1
run kst1 TPGAN_KST=1
run kst1_s12 TPGAN_KST=1 TPGAN_TAP_MAXSTAGES=12
run kst2_s12 TPGAN_TAP_MAXSTAGES=12
run default_again X=1
