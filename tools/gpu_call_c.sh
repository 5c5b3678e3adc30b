#!/bin/bash
mkdir -p gpurun_out
echo "== chain tests" ; timeout 900 python -m pytest tests/test_gpu_chain.py -q -x 2>&1 | tail -15 | tee gpurun_out/c_chain.log
echo "== full-width tests" ; timeout 1200 python -m pytest tests/test_full_width.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/c_full.log
echo "== dynamics + properties" ; timeout 1200 python -m pytest tests/test_gpu_dynamics.py tests/test_gpu_properties.py -q 2>&1 | tail -15 | tee gpurun_out/c_props.log
echo "== fusion sweep pointmaze" ; timeout 600 python tools/fusion_sweep.py pointmaze 1 64 256 512 1024 2048 4096 2>&1 | tail -40 | tee gpurun_out/c_sweep_pm.log
echo "== door units" ; timeout 600 python tools/fusion_sweep.py door 512 4096 2>&1 | tail -40 | tee gpurun_out/c_sweep_door.log
echo "== done"
