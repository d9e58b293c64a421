#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r4e_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r4e_pytest_all.log
for i in 1 2; do
timeout 600 python bench.py --no-cpu --no-secondary > gpurun_out/r4e_bench$i.json 2> gpurun_out/r4e_bench$i.err; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r4e_bench$i.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], d['clocks'])
PY
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r4e_launches_tf32.csv \
      python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu --no-secondary --profile-step > gpurun_out/r4e_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r4e_launches_tf32.csv tf32_b32 | head -14
