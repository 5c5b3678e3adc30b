mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_width.py tests/test_gpu_dynamics.py -m gpu -q -x -k "proj or dyn or full or Door or door or cheetah" > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/m_pytest.log
timeout 300 python tools/projector_paths.py > gpurun_out/m_proj.log 2>&1; cat gpurun_out/m_proj.log | tail -9
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/m_legs.log
import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
pk = bench.peaks()
for name in ('door', 'halfcheetah'):
    r = bench.config_leg(name, dev, pk, {})
    print(name, r['p50_step_latency_ms'], r['unet_tensor_frac_of_sustained'])
PY
