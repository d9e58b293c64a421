#!/bin/bash
# GPU-box script: flat-slab conv correctness + A/B timings (writes gpurun_out/r2h_*)
mkdir -p gpurun_out
python -m pytest tests/test_conv_gpu.py -q -m gpu > gpurun_out/r2h_pytest_conv.log 2>&1; echo "pytest conv rc=$?"; grep -E "FAILED|passed|failed" gpurun_out/r2h_pytest_conv.log | head -20
TPGAN_FLATCONV=0 python -m pytest tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r2h_pytest_model0.log 2>&1; echo "pytest model (noflat) rc=$?"; tail -3 gpurun_out/r2h_pytest_model0.log
python -m pytest tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r2h_pytest_model1.log 2>&1; echo "pytest model (flat auto) rc=$?"; tail -3 gpurun_out/r2h_pytest_model1.log
for cfg in "64 64 1" "128 128 2" "256 256 4" "192 64 1" "384 128 2"; do
  set -- $cfg
  for m in 0 2; do
    echo "== cin $1 cout $2 div $3 FLATCONV=$m"
    TPGAN_FLATCONV=$m timeout 120 python tools/bench_local.py --cin $1 --cout $2 --div $3 --kinds fwd,dgrad,wgrad 2>&1 | tail -3
  done
done > gpurun_out/r2h_local.log 2>&1
cat gpurun_out/r2h_local.log
