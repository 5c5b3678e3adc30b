mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_width.py tests/test_gpu_properties.py -m gpu -q -x -k "fp32 or ill or project or raw or f32" > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/o_pytest.log
timeout 300 python tools/ill_step_cost.py pointmaze 4096 2>&1 | tail -4 | tee gpurun_out/o_ill.log
timeout 300 python tools/ill_step_cost.py halfcheetah 1024 2>&1 | tail -4 | tee -a gpurun_out/o_ill.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/o_f32layers.log
import sys, torch
sys.path.insert(0, '.')
import bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["pointmaze"]; B = 4096; T = 6
net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="fp32", max_batch=B)
dif = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=500)
synthetic.fill_state_dict(dif, 0); dif.to(dev)
eng = dif.engine(32, dev)
x = torch.randn(B, 32, T, device=dev)
eng.unet_forward(x, step=3); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): eng.unet_forward(x, step=3)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("fp32 U-Net pass B=%d: %.2f ms (%.1f TFLOP/s fp32)" % (B, ms, eng.info()["conv_flops_per_sample"] * B / ms / 1e9))
tot = 0
for lay in eng.layers():
    t = eng.time_layer(lay["index"], B, iters=5); tot += t
    print("%-34s L=%2d Cin=%4d Cout=%4d taps=%d  %.3f ms %6.1f TF/s" % (lay["name"], lay["L_out"], lay["C_in"], lay["C_out"], lay["taps"], t, lay["flops_per_sample"] * B / t / 1e9))
print("sum of layers %.2f ms" % tot)
PY
