mkdir -p gpurun_out
timeout 200 python tools/step_times.py 262144 2>&1 | tail -1 | tee gpurun_out/s_stream.log
timeout 200 python tools/step_times.py 65536 2>&1 | tail -1 | tee -a gpurun_out/s_stream.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
timeout 300 python tools/projector_paths.py 4096 2>&1 | tail -1
