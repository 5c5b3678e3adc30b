"""Experiment: does running two independent half-batches concurrently (two streams, two engines) beat one batch?
Tails / set-up / partial rounds of one chain would be filled by the other chain's kernels."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, DynamicsAwarePolicy, synthetic, _native as N

w = bench.WORKLOADS["pointmaze"]
dev = torch.device("cuda", 0)
T, S, H = 6, int(sys.argv[1]) if len(sys.argv) > 1 else 100, 32
P, nz = bench.projector_inputs(w)

def make(B):
    net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="bf16", max_batch=B)
    dif = GaussianDiffusion(net, horizon=H, observation_dim=4, action_dim=2, n_timesteps=S)
    synthetic.fill_state_dict(dif, 0)
    dif.to(dev)
    pol = DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=4, observation_dim=4, action_dim=2,
                              horizon=H, projection_schedule="noise_schedule")
    eng = pol._engine(dev)
    flags = pol._loop_flags(eng) | N.FLAG_CONDITIONS | N.FLAG_PHILOX_INIT
    eng.set_conditions({0: torch.zeros(1, T, device=dev)}, B)
    return eng, flags, torch.empty(B, H, T, device=dev)

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

for parts in (1, 2, 3, 4):
    B = 4096 // parts if 4096 % parts == 0 else (4096 // parts // 32) * 32
    engs = [make(B) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    def run():
        for (eng, flags, x), st in zip(engs, streams):
            with torch.cuda.stream(st):
                eng.sample(x, S, flags=flags, seed=1)
    dt = timeit(run)
    print("parts=%d B_each=%d: %.3f ms per diffusion step of %d plans -> %.0f plans/s at 500 steps" % (
        parts, B, dt / S * 1e3, B * parts, B * parts / (dt / S * 500)))
    del engs
    torch.cuda.empty_cache()
