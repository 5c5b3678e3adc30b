"""Dependency stalls inside the conv chains (needs a -DDAD_TUNING build: DAD_TUNING=1 DAD_LIB_PATH=.../libdad_tune.so):
for every chain of the workload, time at fusion level 2 vs 3 and how many tile waits had to spin.
python tools/chain_stalls.py [workload] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "halfcheetah"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name]["B"]
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS[name])
w["S"] = 20
net, dif = bench.build_policy(w, B, "bf16", dev, latency_max_batch=0)
eng = dif.engine(w["H"], dev)
x = torch.randn(B, w["H"], w["n"] + w["m"], device=dev)
eng.unet_forward(x, step=3)
for level in (2, 3):
    eng.set_fusion(level)
    tot = 0.0
    print("level", level)
    for u in eng.units():
        eng.debug_counters(reset=True)
        ms = eng.time_unit(u["index"], B, iters=10)
        err, spun, ns, waits = eng.debug_counters(reset=True)
        tot += ms
        if u["is_chain"]:
            print("  %-40s %2d convs %8.4f ms %7.1f TF/s   waits %8d  spun %7d (%.1f%%)  spin time %.1f us per launch summed over waiters"
                  % (u["kernel"], u["n_layers"], ms, u["flops_per_sample"] * B / ms / 1e9, waits // 11, spun // 11,
                     100.0 * spun / max(waits, 1), ns / 11 / 1e3))
    print("  sum of units %.4f ms" % tot)
