"""Compare the U-Net output of kernel-path variants (env-selected at engine creation) against the generic
tcgen05 path (DAD_T3=0), at a batch that exercises many tiles.  Diagnostic."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
T = w["n"] + w["m"]
x = torch.randn(B, w["H"], T, device=dev, generator=torch.Generator(device=dev).manual_seed(1))

def run(env):
    for k in ("DAD_T3", "DAD_T3_MODE", "DAD_T3_NS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="bf16", max_batch=max(B, 256))
    dif = GaussianDiffusion(net, horizon=w["H"], observation_dim=w["n"], action_dim=w["m"], n_timesteps=50)
    synthetic.fill_state_dict(dif, 0)
    dif.to(dev)
    out = net(x, torch.full((B,), 7, device=dev, dtype=torch.long))
    torch.cuda.synchronize()
    return out

ref = run({"DAD_T3": "0"})
for env in ({"DAD_T3_MODE": "0"}, {"DAD_T3_MODE": "1"}, {"DAD_T3_MODE": "2", "DAD_T3_NS": "1"}, {"DAD_T3_MODE": "2", "DAD_T3_NS": "2"}):
    out = run(env)
    d = (out - ref).float()
    per_sample = d.flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)
    bad = (per_sample > 2e-2).nonzero().flatten().tolist()
    print(env, "rel-L2 %.3e  max per-sample %.3e  bad samples (%d): %s" % (float(d.norm() / ref.norm()), float(per_sample.max()), len(bad), bad[:24]))
