mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_full_width.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/chk_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/chk_pytest.log
timeout 600 python tools/fusion_sweep.py pointmaze 256 512 1024 2048 4096 2>&1 | grep "B= " | sed 's/L0.*L2:/L2:/' | tee gpurun_out/chk_sweep.log
timeout 600 python tools/fusion_sweep.py door 512 1024 2>&1 | grep "B= " | sed 's/L0.*L2:/L2:/' | tee -a gpurun_out/chk_sweep.log
