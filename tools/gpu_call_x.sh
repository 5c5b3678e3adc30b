mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -k "lean or tensor_core_projector or fp32_sibling" > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/x_pytest.log
