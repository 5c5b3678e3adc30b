mkdir -p gpurun_out
timeout 300 python tools/compare_libs.py 600 libdad_p0.so libdad_p1.so 2>&1 | tail -3 | tee gpurun_out/u_cmp.log
for lib in libdad_p0 libdad_p1; do
  echo "== $lib stalls"
  DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/$lib.so timeout 300 python tools/chain_stalls.py pointmaze 4096 2>&1 | grep -A8 "level 3" | tee gpurun_out/u_stalls_$lib.log
done
for lib in libdad_p0 libdad_p1 libdad_p0 libdad_p1; do
  echo "== $lib"
  DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/$lib.so timeout 600 python tools/fusion_sweep.py pointmaze 512 1024 4096 2>&1 | grep "B= " | cut -c1-20,150-260 | tee -a gpurun_out/u_$lib.log
done
for lib in libdad_p0 libdad_p1; do
  DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/$lib.so timeout 600 python tools/fusion_sweep.py halfcheetah 1024 2>&1 | grep "B= " | cut -c1-20,150-260 | tee -a gpurun_out/u_hc_$lib.log
done
