"""Short driver for ncu: a few U-Net forwards + fused step kernels of the headline workload (PointMaze, B=4096)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, DynamicsAwarePolicy, synthetic, _native as N  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[name]
B = int(sys.argv[3]) if len(sys.argv) > 3 else w["B"]
dev = torch.device("cuda", 0)
T = w["n"] + w["m"]
net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="bf16", max_batch=B)
dif = GaussianDiffusion(net, horizon=w["H"], observation_dim=w["n"], action_dim=w["m"], n_timesteps=w["S"])
synthetic.fill_state_dict(dif, 0)
dif.to(dev)
P, nz = bench.projector_inputs(w)
pol = DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=w["n"], observation_dim=w["n"],
                          action_dim=w["m"], horizon=w["H"], projection_schedule="noise_schedule")
eng = pol._engine(dev)
flags = pol._loop_flags(eng)
x = torch.randn(B, w["H"], T, device=dev)
for r in range(reps):
    eps = eng.unet_forward(x, step=w["S"] // 2)
    eng.step(x, eps, w["S"] // 2, flags=flags, seed=1)
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
