#!/bin/bash
mkdir -p gpurun_out
for lib in libdad_b200 libdad_n3c16 libdad_n4c16; do
  echo "== $lib"
  DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/$lib.so timeout 600 python tools/fusion_sweep.py pointmaze 512 4096 2>&1 | tail -17 | tee gpurun_out/i_$lib.log
done
echo "== done"
