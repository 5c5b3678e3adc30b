mkdir -p gpurun_out
timeout 600 python tools/ill_math_modes.py 2>&1 | grep -v Warning | tee gpurun_out/p_modes.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_width.py -m gpu -q -x > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log
timeout 300 python tools/ill_step_cost.py pointmaze 4096 2>&1 | tail -4 | tee gpurun_out/p_ill.log
