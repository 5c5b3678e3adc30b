"""A small sampling workload that touches every kernel family (for `compute-sanitizer --tool memcheck|racecheck|synccheck`):
chains (ragged batch), generic tcgen05 convs (weight-stationary), latency kernels, the TF32 sibling + warp GroupNorm, the
lean and the generic step kernels, fused and tensor-core projectors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamics_aware_diffusion_b200 import (TemporalUnet, GaussianDiffusion, GuidedPolicy, DynamicsAwarePolicy,
                                           ProjectionMatrixBuilder, synthetic)

dev = torch.device("cuda", 0)
A, Bm = synthetic.double_integrator(0.1)
P = ProjectionMatrixBuilder(A, Bm, 4, 2).get_projection_matrix(32)
nz = synthetic.SyntheticNormalizer(4, 2)
start = torch.zeros(1, 6, device=dev)
for dim, mults, lat in ((128, (1, 2, 4), 0), (64, (1, 2), 8)):
    net = TemporalUnet(6, dim=dim, dim_mults=mults, precision="bf16", max_batch=300, latency_max_batch=lat)
    dif = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=3)
    synthetic.fill_state_dict(dif, 1)
    dif.to(dev)
    for pol in (GuidedPolicy(dif, nz), DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=4, observation_dim=4,
                                                           action_dim=2, horizon=32, projection_schedule="noise_schedule")):
        for B in (3, 300):
            x = pol.sample_loop(batch_size=B, conditions={0: start}, seed=1)
            y, tr = pol.sample_loop(batch_size=B, conditions={0: start}, seed=1, return_trace=True)
            torch.cuda.synchronize()
            assert bool(torch.isfinite(x).all()) and torch.equal(x, y)
    print("ok", dim, mults, flush=True)
