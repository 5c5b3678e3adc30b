#!/bin/bash
# Round-2 GPU session A: chain-kernel correctness first (tight timeouts), then the whole GPU suite, smoke, bench, sweeps.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
echo "== chain tests" ; timeout 900 python -m pytest tests/test_gpu_chain.py -q -x 2>&1 | tail -25 | tee gpurun_out/a_chain.log
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -12 | tee gpurun_out/a_smoke.log
echo "== fusion sweep pointmaze" ; timeout 600 python tools/fusion_sweep.py pointmaze 64 512 1024 4096 2>&1 | tail -40 | tee gpurun_out/a_sweep_pm.log
echo "== full-width tests" ; timeout 1200 python -m pytest tests/test_full_width.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/a_full.log
echo "== gpu suite" ; timeout 2400 python -m pytest tests -q -m gpu --deselect tests/test_full_width.py --deselect tests/test_gpu_chain.py 2>&1 | tail -25 | tee gpurun_out/a_suite.log
echo "== bench" ; timeout 1500 python bench.py --steps 5 --warmup 3 --layers-out gpurun_out/a_layers.json > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
echo "== fusion sweep halfcheetah/door" ; timeout 600 python tools/fusion_sweep.py halfcheetah 1024 2>&1 | tail -30 | tee gpurun_out/a_sweep_hc.log
timeout 600 python tools/fusion_sweep.py door 4096 2>&1 | tail -30 | tee gpurun_out/a_sweep_door.log
echo "== done"
