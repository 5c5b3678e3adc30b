#!/bin/bash
mkdir -p gpurun_out
echo "== default lib B=4096"; timeout 300 python tools/chain_diag.py pointmaze 4096 2>&1 | tee gpurun_out/b_diag_4096.log | tail -50
echo "== default lib B=64"; timeout 300 python tools/chain_diag.py pointmaze 64 2>&1 | tee gpurun_out/b_diag_64.log | tail -50
echo "== small-params lib B=4096"; DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/libdad_mc4.so timeout 300 python tools/chain_diag.py pointmaze 4096 2>&1 | tee gpurun_out/b_diag_mc4_4096.log | tail -50
echo "== small-params lib B=64"; DAD_TUNING=1 DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/libdad_mc4.so timeout 300 python tools/chain_diag.py pointmaze 64 2>&1 | tee gpurun_out/b_diag_mc4_64.log | tail -50
echo "== new gpu tests"; timeout 600 python -m pytest tests/test_gpu_dynamics.py -q 2>&1 | tail -15 | tee gpurun_out/b_dyn.log
echo "== ncu chain L=32 (first chain launch of a pass) and C=512 chain"
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_chain_kernel -s 12 -c 3 -f -o gpurun_out/b_chain python /tmp/prof_unet.py > gpurun_out/b_ncu.log 2>&1
ls -la gpurun_out/b_chain* 
echo "== done"
