"""Same box, same process: in-loop p50 step latency (graph replays) AND the sum of the isolated unit times for every
fusion level, with the SM clock sampled during each in-loop measurement.  python tools/chain_diag2.py [workload] [B]"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from dynamics_aware_diffusion_b200 import _native as N  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS[name])
w["S"] = 200
net, dif = bench.build_policy(w, B, "bf16", dev, latency_max_batch=0)
pol, eng, flags, start, dyn = bench.attach_policy(dif, w, dev, B)
x = torch.empty(B, w["H"], w["n"] + w["m"], device=dev)
print("workload %s B=%d" % (name, B))
for rep in range(reps):
    for level in (0, 1, 2, 3):
        eng.set_fusion(level)
        eng.sample_profile(x, 20, flags=flags | N.FLAG_PHILOX_INIT, seed=1)
        cs = bench.ClockSampler(0)
        cs.start()
        t0 = time.perf_counter()
        ms = eng.sample_profile(x, 200, flags=flags | N.FLAG_PHILOX_INIT, seed=2)
        wall = time.perf_counter() - t0
        clk = cs.finish()
        units = eng.units()
        iso = sum(eng.time_unit(u["index"], B, iters=20) for u in units)
        print("rep %d level %d: in-loop p50 %.4f ms mean %.4f (wall/200 %.4f)  clock %s MHz %s   isolated unit sum %.4f ms (%d units)"
              % (rep, level, statistics.median(ms), sum(ms) / len(ms), wall / 200 * 1e3, clk["sm_mhz"], clk["reasons"], iso, len(units)), flush=True)
