#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_conv_bf16_gpu.py -x -q -m gpu -k "wgrad or adjoint" > gpurun_out/r4i_pytest_wgrad.log 2>&1; echo "pytest wgrad rc=$?"; tail -5 gpurun_out/r4i_pytest_wgrad.log
for sh in enh128 enh64 enh32 enh16 conv4rb; do for wp in 0 1; do TPGAN_WGRAD_PAIR=$wp python tools/bench_conv.py --only $sh --kinds wgrad | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  wpair=$wp %-22s %-5s %.4f ms %6.1f TF' % (d['shape'], d['kind'], d['ms'], d['tflops']))"; done; done
for wp in 0 1 0 1; do
TPGAN_WGRAD_PAIR=$wp timeout 600 python bench.py --no-cpu --no-secondary 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('wpair=$wp bench', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))"
done
