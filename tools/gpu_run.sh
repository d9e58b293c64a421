#!/bin/bash
mkdir -p gpurun_out
for fl in 0 1 0 1; do
  TPGAN_FLATCONV=$fl python bench.py --no-cpu --no-secondary > gpurun_out/r2r_bench_fl$fl.json 2> gpurun_out/r2r_bench.err; echo "bench flat=$fl rc=$?"
  python - <<EOF
import json
for l in open("gpurun_out/r2r_bench_fl$fl.json"):
    if l.startswith("{"):
        d=json.loads(l); print("flat=$fl", d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
EOF
done
python -m pytest tests -x -q -m gpu > gpurun_out/r2r_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2r_pytest_all.log
