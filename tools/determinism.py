"""Run the U-Net forward several times on the same input and report bitwise differences (race detector)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
T = w["n"] + w["m"]
net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="bf16", max_batch=B)
dif = GaussianDiffusion(net, horizon=w["H"], observation_dim=w["n"], action_dim=w["m"], n_timesteps=50)
synthetic.fill_state_dict(dif, 0)
dif.to(dev)
x = torch.randn(B, w["H"], T, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
t = torch.full((B,), 7, device=dev, dtype=torch.long)
ref = net(x, t).clone()
bad = 0
for r in range(reps):
    out = net(x, t)
    d = (out != ref)
    if bool(d.any()):
        bad += 1
        rows = d.flatten(1).any(dim=1).nonzero().flatten()
        print("rep %d: %d differing samples, first %s, max abs diff %.3e" % (r, rows.numel(), rows[:12].tolist(), float((out - ref).abs().max())))
print("env", {k: v for k, v in os.environ.items() if k.startswith("DAD_")}, "nondeterministic reps:", bad, "/", reps)
