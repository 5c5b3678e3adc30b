mkdir -p gpurun_out
timeout 300 python tools/compare_libs.py 600 libdad_p0.so libdad_p1.so 2>&1 | tail -4 | tee gpurun_out/t_cmp.log
for lib in libdad_p0 libdad_p1 libdad_p0 libdad_p1; do
  echo "== $lib"
  DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/$lib.so timeout 600 python tools/fusion_sweep.py pointmaze 512 4096 2>&1 | tail -16 | tee -a gpurun_out/t_$lib.log
done
DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/libdad_p0.so timeout 600 python tools/fusion_sweep.py halfcheetah 1024 2>&1 | tail -9 | tee gpurun_out/t_hc_p0.log
DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/libdad_p1.so timeout 600 python tools/fusion_sweep.py halfcheetah 1024 2>&1 | tail -9 | tee gpurun_out/t_hc_p1.log
