"""Per-step wall time of the value-guided loop (ValueGuidedPolicy.sample_loop) vs the unguided graph-replayed loop."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, GuidedPolicy, ValueGuidedPolicy, synthetic

dev = torch.device("cuda", 0)
S = 100
net = TemporalUnet(6, dim=128, dim_mults=(1, 2, 4), precision="bf16", max_batch=64)
dif = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=S)
synthetic.fill_state_dict(dif, 0)
dif.to(dev)
nz = synthetic.SyntheticNormalizer(4, 2)
value = torch.nn.Sequential(torch.nn.Linear(4, 64), torch.nn.Mish(), torch.nn.Linear(64, 1)).to(dev)
pols = {"unguided": GuidedPolicy(dif, nz), "value-guided": ValueGuidedPolicy(dif, nz, value, guide_weight=0.1)}
start = torch.zeros(1, 6, device=dev)
for name, pol in pols.items():
    for B in (1, 8, 64):
        ts = []
        for k in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = pol.sample_loop(batch_size=B, conditions={0: start}, seed=k)
            _ = out[0, :2].cpu()
            ts.append((time.perf_counter() - t0) * 1e3)
        print("%s B=%d: %.1f us/step" % (name, B, statistics.median(ts[1:]) * 1e3 / S), flush=True)
