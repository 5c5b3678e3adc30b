#!/bin/bash
mkdir -p gpurun_out
echo "== chain + full tests" ; timeout 1500 python -m pytest tests/test_gpu_chain.py tests/test_full_width.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/g_tests.log
echo "== bench full" ; timeout 1800 python bench.py --steps 5 --warmup 3 --layers-out gpurun_out/g_layers.json > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; tail -3 gpurun_out/g_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/g_bench.json'))
for k in ('value','ms_per_step','e2e','p50_step_latency_ms','unet_tensor_frac_of_sustained','clocks','plan_latency_b1_ms','gpu_eager_baseline','cpu_baseline'): print(k, d.get(k))
print(d['roofline'])
for c in d.get('configs',[]): print(c)
PY
echo "== done"
