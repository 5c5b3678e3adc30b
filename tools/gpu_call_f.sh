#!/bin/bash
mkdir -p gpurun_out
export DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/libdad_tune.so
echo "== stalls halfcheetah" ; timeout 600 python tools/chain_stalls.py halfcheetah 1024 2>&1 | tail -40 | tee gpurun_out/f_stalls_hc.log
echo "== stalls pointmaze" ; timeout 600 python tools/chain_stalls.py pointmaze 4096 2>&1 | tail -40 | tee gpurun_out/f_stalls_pm.log
echo "== stalls pointmaze 512" ; timeout 600 python tools/chain_stalls.py pointmaze 512 2>&1 | tail -40 | tee gpurun_out/f_stalls_pm512.log
echo "== done"
