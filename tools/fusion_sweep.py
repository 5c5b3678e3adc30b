"""p50 diffusion-step latency per fusion level (dad_set_fusion) and batch: what the conv chains buy over the per-layer
kernels of round 1.  python tools/fusion_sweep.py [pointmaze|halfcheetah|door] [B ...]"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from dynamics_aware_diffusion_b200 import _native as N  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
batches = [int(v) for v in sys.argv[2:]] or [64, 512, 1024, 4096]
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS[name])
w["S"] = 40
net, dif = bench.build_policy(w, max(batches), "bf16", dev, latency_max_batch=0)
pol, eng, flags, start, dyn = bench.attach_policy(dif, w, dev, max(batches))
info = eng.info()
print("%s: %d conv layers, %.1f MFLOP per sample and step" % (name, info["n_conv_layers"], info["conv_flops_per_sample"] / 1e6))
for B in batches:
    x = torch.empty(B, w["H"], w["n"] + w["m"], device=dev)
    eng.set_conditions({0: start}, B)
    row = []
    for level in (0, 1, 2, 3):
        eng.set_fusion(level)
        eng.sample_profile(x, 5, flags=flags | N.FLAG_PHILOX_INIT, seed=1)
        ms = statistics.median(eng.sample_profile(x, 30, flags=flags | N.FLAG_PHILOX_INIT, seed=2))
        row.append((level, ms, eng.info()["launches_per_step"]))
    print("B=%6d  " % B + "   ".join("L%d: %.4f ms (%d launches, %.0f TF/s)" % (lv, ms, nl, info["conv_flops_per_sample"] * B / ms / 1e9)
                                      for lv, ms, nl in row), flush=True)
eng.set_fusion(3)
print("units at level 3:")
for u in eng.units():
    t = eng.time_unit(u["index"], batches[-1], iters=10)
    print("  %-44s %2d layers  %8.4f ms  %7.1f TF/s" % (u["kernel"], u["n_layers"], t, u["flops_per_sample"] * batches[-1] / t / 1e9))
