#!/bin/bash
mkdir -p gpurun_out
echo "== chain tests" ; timeout 900 python -m pytest tests/test_gpu_chain.py -q -x 2>&1 | tail -5 | tee gpurun_out/d_chain.log
echo "== fusion sweep pointmaze" ; timeout 600 python tools/fusion_sweep.py pointmaze 64 512 2048 4096 2>&1 | tail -40 | tee gpurun_out/d_sweep_pm.log
echo "== halfcheetah" ; timeout 600 python tools/fusion_sweep.py halfcheetah 1024 2>&1 | head -3 | tee gpurun_out/d_sweep_hc.log
echo "== diag2" ; timeout 600 python tools/chain_diag2.py pointmaze 4096 1 2>&1 | tail -6 | tee gpurun_out/d_diag2.log
echo "== done"
