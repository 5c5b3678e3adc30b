"""Plan latency of GuidedPolicy.sample_loop at small batch (the get_action shape): latency kernels vs throughput kernels."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, DynamicsAwarePolicy, synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
T = w["n"] + w["m"]
P, nz = bench.projector_inputs(w)
for lat in (0, 8):
    net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="bf16", max_batch=8, latency_max_batch=lat)
    dif = GaussianDiffusion(net, horizon=w["H"], observation_dim=w["n"], action_dim=w["m"], n_timesteps=w["S"])
    synthetic.fill_state_dict(dif, 0)
    dif.to(dev)
    pol = DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=w["n"], observation_dim=w["n"],
                              action_dim=w["m"], horizon=w["H"], projection_schedule="noise_schedule", projection_strength=1.0)
    start = torch.zeros(1, T, device=dev)
    for B in (1, 4):
        ts = []
        for k in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = pol.sample_loop(batch_size=B, conditions={0: start}, seed=k)
            _ = out[0, :2].cpu()
            ts.append((time.perf_counter() - t0) * 1e3)
        print("%s latency_max_batch=%d B=%d: plan %.1f ms (%d steps, %.1f us/step)" % (
            name, lat, B, statistics.median(ts[1:]), w["S"], statistics.median(ts[1:]) * 1e3 / w["S"]), flush=True)
    del pol, dif, net
    torch.cuda.empty_cache()
