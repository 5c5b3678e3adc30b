# Final single-GPU session of the round: the driver's own sequence (GPU tests, smoke, reference arm, bench).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke gpurun_out/final_smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','p50_step_latency_ms','unet_tensor_frac_of_sustained','clocks','gpu_launches','cpu_baseline'): print(k, d.get(k))
print(d['roofline']['frac'], d['roofline_step_kernel_stream']['philox']['frac'], d['roofline_step_kernel_stream']['injected_noise']['frac'])
PY
