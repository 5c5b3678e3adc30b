"""Cost of `fp32_ill_conditioned_steps`: one fp32 U-Net pass (SIMT kernels) per sampling loop vs the loop itself."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name]["B"]
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS[name])
net, dif = bench.build_policy(w, B, "bf16", dev)
pol, eng, flags, start, dyn = bench.attach_policy(dif, w, dev, B)
for on in (False, True, False, True):
    dif.fp32_ill_conditioned_steps = on
    pol.sample_loop(batch_size=B, conditions={0: start}, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x = pol.sample_loop(batch_size=B, conditions={0: start}, seed=2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s B=%d S=%d fp32_ill_conditioned_steps=%s: %.1f ms per loop (%.0f plans/s), finite=%s" %
          (name, B, w["S"], on, dt * 1e3, B / dt, bool(torch.isfinite(x).all())), flush=True)
