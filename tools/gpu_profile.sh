#!/bin/bash
# GPU-box script: ncu --set full captures of the round's kernel cases (one launch each), after a plain run of the same command.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python tools/bench_conv.py --only enh128 --kinds wgrad --iters 1 > /dev/null 2>&1 && \
$NCU -k regex:wgrad_kernel -s 2 -c 1 -o gpurun_out/ncu_full_r2_wgrad_enh128 python tools/bench_conv.py --only enh128 --kinds wgrad --iters 1 > gpurun_out/ncu_r2_1.log 2>&1; echo "wgrad enh128 rc=$?"
python tools/bench_conv.py --only add128 --kinds wgrad --iters 1 > /dev/null 2>&1 && \
$NCU -k regex:wgrad_kernel -s 2 -c 1 -o gpurun_out/ncu_full_r2_wgrad_add128 python tools/bench_conv.py --only add128 --kinds wgrad --iters 1 > gpurun_out/ncu_r2_2.log 2>&1; echo "wgrad add128 rc=$?"
python tools/bench_local.py --cin 128 --cout 128 --div 2 --kinds fwd --iters 1 > /dev/null 2>&1 && \
TPGAN_FLATCONV=0 $NCU -k regex:tapgemm_kernel -s 2 -c 1 -o gpurun_out/ncu_full_r2_local_tapgemm python tools/bench_local.py --cin 128 --cout 128 --div 2 --kinds fwd --iters 1 > gpurun_out/ncu_r2_3.log 2>&1; echo "local tapgemm rc=$?"
TPGAN_FLATCONV=2 $NCU -k regex:flatconv_kernel -s 2 -c 1 -o gpurun_out/ncu_full_r2_local_flatconv python tools/bench_local.py --cin 128 --cout 128 --div 2 --kinds fwd --iters 1 > gpurun_out/ncu_r2_4.log 2>&1; echo "local flatconv rc=$?"
$NCU -k regex:wgrad_kernel -s 2 -c 1 -o gpurun_out/ncu_full_r2_local_wgrad python tools/bench_local.py --cin 128 --cout 128 --div 2 --kinds wgrad --iters 1 > gpurun_out/ncu_r2_5.log 2>&1; echo "local wgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
