#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/prof_unet.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
w = dict(bench.WORKLOADS['pointmaze']); w['S'] = 20
net, dif = bench.build_policy(w, 4096, 'bf16', dev, latency_max_batch=0)
eng = dif.engine(32, dev)
x = torch.randn(4096, 32, 6, device=dev)
for _ in range(3):
    eng.unet_forward(x, step=3)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_chain_kernel -s 12 -c 3 -f -o gpurun_out/e_chain python /tmp/prof_unet.py > gpurun_out/e_ncu.log 2>&1
ls -la gpurun_out/e_chain*
echo "== gpu suite" ; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/e_suite.log
echo "== bench" ; timeout 1500 python bench.py --steps 5 --warmup 3 --no-extra-legs --layers-out gpurun_out/e_layers.json > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/e_bench.json'))
for k in ('value','ms_per_step','e2e','p50_step_latency_ms','unet_tensor_frac_of_sustained','roofline','clocks','plan_latency_b1_ms'): print(k, d.get(k))
PY
echo "== done"
