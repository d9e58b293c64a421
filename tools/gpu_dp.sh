#!/bin/bash
# GPU-box script (N GPUs): one data-parallel bench line, N = $1
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-secondary > gpurun_out/dp${N}.json 2> gpurun_out/dp${N}.err; echo "dp$N rc=$?"
python - <<PY
import json
for l in open('gpurun_out/dp${N}.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['n_gpus'], d.get('replica_checksum_spread'), d['clocks'])
PY
