"""Accuracy and cost of the fp32 sibling's arithmetic (GaussianDiffusion.ill_conditioned_math) at the ill-conditioned step:
eps rel-L2 vs the reference, the teacher-forced x error of that step, and the time of one U-Net pass at B=4096."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
import helpers
import test_gpu_parity as T
import bench

dev = torch.device("cuda", 0)
for name in ("tiny", "pointmaze", "door_s"):
    c, g, dif, sd = T.models(name, "bf16")
    S = c["S"]
    k_eps = list(g["unet_steps"]).index(S - 1)
    x = T.cu(g["x_init"])
    for mode in ("bf16", "tf32", "tf32x3", "fp32"):
        if mode == "bf16":
            dif.fp32_ill_conditioned_steps = False
        else:
            dif.fp32_ill_conditioned_steps = True
            dif.ill_conditioned_math = mode
        eng = dif.engine(c["H"], dev)
        e = dif.eps_engine(eng, S - 1)
        eps = e.unet_forward(x, step=S - 1)
        err_eps = helpers.rel_l2(eps.cpu().numpy(), g["unet_eps"][k_eps])
        xs = x.clone()
        eng.step(xs, eps, S - 1, noise=T.cu(g["noise"][0]))
        err_x = helpers.rel_l2(xs.cpu().numpy(), g["trace_plain"][0])
        print("%-10s %-7s eps rel-L2 %.3e   x after the ill-conditioned step rel-L2 %.3e" % (name, mode, err_eps, err_x), flush=True)
    dif.fp32_ill_conditioned_steps, dif.ill_conditioned_math = True, "tf32"

w = bench.WORKLOADS["pointmaze"]
B = 4096
net, dif = bench.build_policy(w, B, "bf16", dev)
x = torch.randn(B, 32, 6, device=dev)
for mode in ("fp32", "tf32", "tf32x3"):
    dif.ill_conditioned_math = mode
    eng = dif.engine(32, dev)
    e = dif.eps_engine(eng, w["S"] - 1)
    e.unet_forward(x, step=w["S"] - 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        e.unet_forward(x, step=w["S"] - 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("U-Net pass of the fp32 sibling, B=%d, %-7s: %.2f ms (%.0f TFLOP/s)" % (B, mode, ms, eng.info()["conv_flops_per_sample"] * B / ms / 1e9), flush=True)
