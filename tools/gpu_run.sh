#!/bin/bash
# GPU-box script (scratch): writes gpurun_out/r2k_*
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2k_pytest_all.log
for only in fwd dgrad single grouped; do
  echo "== FLAT_ONLY=$only"
  TPGAN_FLATCONV=1 TPGAN_FLAT_ONLY=$only python -m pytest tests/test_model_gpu.py -x -q -m gpu -k batch32 2>&1 | grep -E "AssertionError: [0-9]|passed|failed" | head -3
done
run() { echo "== $*"; env "$@" timeout 120 python tools/bench_local.py --cin $CIN --cout $COUT --div $DIV --kinds wgrad 2>&1 | grep -E "wgrad\[0|kind" | sort -u | cut -c1-330; }
CIN=128 COUT=128 DIV=2
run TPGAN_WGRAD_DEBUG=1
run TPGAN_WGRAD_DEBUG=1 TPGAN_WGRAD_TAPPACK=4
run TPGAN_WGRAD_DEBUG=1 TPGAN_WGRAD_TAPPACK=4 TPGAN_WGRAD_PX=64
CIN=256 COUT=256 DIV=4
run TPGAN_WGRAD_DEBUG=1
run TPGAN_WGRAD_DEBUG=1 TPGAN_WGRAD_MPU_MINPIX=1000
CIN=512 COUT=512 DIV=8
run TPGAN_WGRAD_DEBUG=1
run TPGAN_WGRAD_DEBUG=1 TPGAN_WGRAD_MPU_MINPIX=100
python bench.py --no-cpu --no-secondary --per-layer gpurun_out/r2k_per_layer.jsonl > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<EOF
import json
for l in open("gpurun_out/r2k_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]); r=d["roofline"]; print(r["kernel"][:20], r["frac"], r["share_of_step"], {k:(round(v["frac"],3), round(v["share_of_step"],3)) for k,v in r["other_kernels"].items()})
EOF
