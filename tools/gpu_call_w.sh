mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/w_pytest.log
timeout 600 python tools/ill_math_modes.py 2>&1 | grep -v Warning | tail -16 | tee gpurun_out/w_modes.log
timeout 300 python tools/ill_step_cost.py pointmaze 4096 2>&1 | tail -4 | tee gpurun_out/w_ill.log
