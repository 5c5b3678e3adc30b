"""Per-layer timing table of a workload (uses the C-ABI measurement hooks); DAD_TC_DEBUG=1|2 isolates
mainloop / epilogue."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "pointmaze"
w = bench.WORKLOADS[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else w["B"]
dev = torch.device("cuda", 0)
T = w["n"] + w["m"]
net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="bf16", max_batch=B)
dif = GaussianDiffusion(net, horizon=w["H"], observation_dim=w["n"], action_dim=w["m"], n_timesteps=w["S"])
synthetic.fill_state_dict(dif, 0)
dif.to(dev)
eng = dif.engine(w["H"], dev)
x = torch.randn(B, w["H"], T, device=dev)
eng.unet_forward(x, step=3)
tot = 0.0
rows = []
for lay in eng.layers():
    ms = eng.time_layer(lay["index"], B, iters=20)
    tot += ms
    rows.append((lay, ms))
print("mode", os.environ.get("DAD_TC_DEBUG", "0"), "total %.4f ms" % tot)
for lay, ms in rows:
    print("%-34s L=%2d Cin=%4d Cout=%4d taps=%d  %.4f ms %7.1f TF/s" % (lay["name"], lay["L_out"], lay["C_in"], lay["C_out"], lay["taps"], ms, lay["flops_per_sample"] * B / ms / 1e9))
