"""Per-layer time at fusion level 0 (conv_t3 per layer) vs level 1 (conv_chain per layer), and per-unit times at levels
2 / 3, for one batch: where the chain kernel loses or gains.  python tools/chain_diag.py [workload] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS[name])
w["S"] = 20
net, dif = bench.build_policy(w, B, "bf16", dev, latency_max_batch=0)
eng = dif.engine(w["H"], dev)
x = torch.randn(B, w["H"], w["n"] + w["m"], device=dev)
eng.unet_forward(x, step=3)
per = {}
for level in (0, 1):
    eng.set_fusion(level)
    per[level] = [eng.time_unit(u["index"], B, iters=20) for u in eng.units()]
layers = eng.layers()
print("lib", os.environ.get("DAD_LIB_PATH", "default"), "B", B)
print("%-36s %3s %5s %5s %4s  %9s %9s %7s" % ("layer", "L", "Cin", "Cout", "taps", "L0 us", "L1 us", "diff"))
for lay, t0, t1 in zip(layers, per[0], per[1]):
    print("%-36s %3d %5d %5d %4d  %9.2f %9.2f %+7.2f  %s" % (lay["name"], lay["L_out"], lay["C_in"], lay["C_out"], lay["taps"],
                                                               t0 * 1e3, t1 * 1e3, (t1 - t0) * 1e3, lay["kernel"]))
print("sum L0 %.1f us, L1 %.1f us" % (sum(per[0]) * 1e3, sum(per[1]) * 1e3))
for level in (2, 3):
    eng.set_fusion(level)
    units = eng.units()
    ts = [eng.time_unit(u["index"], B, iters=20) for u in units]
    print("level %d: sum %.1f us: " % (level, sum(ts) * 1e3) + " ".join("%d:%.1f" % (u["n_layers"], t * 1e3) for u, t in zip(units, ts)))
