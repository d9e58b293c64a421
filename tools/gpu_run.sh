#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r3c_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r3c_pytest_all.log
python bench.py --no-cpu --no-secondary --per-layer gpurun_out/r3c_per_layer.jsonl > gpurun_out/r3c_bench.json 2>/dev/null
python - <<EOF
import json
for l in open('gpurun_out/r3c_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
for l in open('gpurun_out/r3c_per_layer.jsonl'):
    r=json.loads(l)
    if r['kind']=='wgrad' and r['tflops']<60 and r['ms']>0.025: print(r['ms'], r['tflops'], r['label'][:70])
EOF
