#!/bin/bash
mkdir -p gpurun_out
for pm in 1 0 1; do
TPGAN_PAIR=$pm timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-secondary > gpurun_out/r4g_dp2_pair$pm.json 2> gpurun_out/r4g_dp2_pair$pm.err; echo "dp2 pair=$pm rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r4g_dp2_pair$pm.json'):
    if l.startswith('{'):
        d=json.loads(l); print('pair=$pm', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['n_gpus'], d.get('replica_checksum_spread'), d['clocks'])
PY
done
timeout 300 python bench.py --no-cpu --no-secondary | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n1', round(d['value'],1), round(d['ms_per_step'],3))"
