#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r4b_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r4b_pytest_all.log
for pm in 0 1; do
TPGAN_PAIR=$pm timeout 600 python bench.py --no-cpu --no-secondary --dtype bf16 > gpurun_out/r4b_bf16_pair$pm.json 2> gpurun_out/r4b_bf16_pair$pm.err; echo "bench bf16 pair=$pm rc=$?"
TPGAN_PAIR=$pm timeout 600 python bench.py --no-cpu --workload pretrain > gpurun_out/r4b_pretrain_pair$pm.json 2> gpurun_out/r4b_pretrain_pair$pm.err; echo "bench pretrain pair=$pm rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r4b_*_pair*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'])
PY
