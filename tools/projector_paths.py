"""Dynamics-projector step at D = 192 (PointMaze H=32): fused SIMT kernel vs pointwise + bf16x3 tensor-core GEMM, per batch.
Decides kProjTcMinBatch (csrc/dad_api.cu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion

dev = torch.device("cuda", 0)
H, T = 32, 6
Bmax = 262144
net = TemporalUnet(T, dim=64, dim_mults=(1,), precision="bf16", max_batch=Bmax)
dif = GaussianDiffusion(net, horizon=H, observation_dim=4, action_dim=2, n_timesteps=100).to(dev)
dif.fp32_ill_conditioned_steps = False      # only the step kernels run here
pol, eng, flags, _, _ = bench.attach_policy(dif, dict(bench.WORKLOADS["pointmaze"], S=100), dev, Bmax)
for B in [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192, 16384, 65536, 262144]:
    row = []
    for extra in (0x400, 0x200):
        eng.time_step_kernel(B, 50, flags=flags | extra, iters=3)
        row.append(eng.time_step_kernel(B, 50, flags=flags | extra, iters=20) * 1e3)
    print("B=%7d  fused SIMT %9.1f us   pointwise + tcgen05 GEMM %9.1f us   (%.0f / %.0f GB/s algorithmic)" % (
        B, row[0], row[1], 12 * H * T * B / row[0] / 1e3, 12 * H * T * B / row[1] / 1e3))
