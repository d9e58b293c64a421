#!/bin/bash
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-secondary"
show() { python -c "
import json,sys
for l in open('$1'):
    if l.startswith('{'):
        d=json.loads(l); print('$2', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d.get('replica_checksum_spread'))
"; }
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu --no-secondary > gpurun_out/r2x_n1.json 2>/dev/null; show gpurun_out/r2x_n1.json n1
i=0
for v in "0 32" "-1 96" "0 96" "-1 64" "-1 160" "-1 96" "0 32"; do
  set -- $v; i=$((i+1))
  TPGAN_COMM_PRIORITY=$1 $T --bucket-mb $2 > gpurun_out/r2x_$i.json 2>gpurun_out/r2x_$i.err; show gpurun_out/r2x_$i.json "dp2 prio=$1 bucket=$2"
done
