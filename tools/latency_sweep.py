import os, sys, time, statistics
sys.path.insert(0, "/root/repo")
import torch, bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, GuidedPolicy, synthetic
w = bench.WORKLOADS["pointmaze"]; dev = torch.device("cuda", 0); T = 6
for lat in (0, 64):
    net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="bf16", max_batch=64, latency_max_batch=lat)
    dif = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=100)
    synthetic.fill_state_dict(dif, 0); dif.to(dev)
    pol = GuidedPolicy(dif, synthetic.SyntheticNormalizer(4, 2))
    start = torch.zeros(1, T, device=dev)
    for B in (1, 8, 12, 16, 24, 32, 48, 64):
        ts = []
        for k in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = pol.sample_loop(batch_size=B, conditions={0: start}, seed=k); _ = out[0, :2].cpu()
            ts.append((time.perf_counter() - t0) * 1e3)
        print("lat=%d B=%d: %.1f us/step" % (lat, B, statistics.median(ts[1:]) * 1e3 / 100), flush=True)
