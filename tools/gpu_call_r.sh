mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r_pytest.log
timeout 200 python tools/step_times.py 262144 2>&1 | tail -1 | tee gpurun_out/r_stream.log
timeout 200 python tools/step_times.py 65536 2>&1 | tail -1 | tee -a gpurun_out/r_stream.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','p50_step_latency_ms','clocks','gpu_launches','roofline_step_kernel_stream'): print(k, d.get(k))
PY
