mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py tests/test_full_width.py -m gpu -q -x > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
timeout 300 python tools/projector_paths.py > gpurun_out/l_proj.log 2>&1; cat gpurun_out/l_proj.log | tail -9
for ws in 0 1; do
  DAD_TUNING=1 DAD_TC_WS=$ws DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/libdad_tune.so timeout 300 python tools/layer_times.py pointmaze 4096 2>&1 | grep -i "downs.*conv\|ups.*conv\|total\|final\|Downsample\|Upsample\|taps=[234] " > gpurun_out/l_layers_ws$ws.log; cat gpurun_out/l_layers_ws$ws.log
done
DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/libdad_tune.so timeout 300 python tools/layer_times.py halfcheetah 1024 2>&1 | grep "total\|taps=[234] " > gpurun_out/l_layers_hc.log; cat gpurun_out/l_layers_hc.log
