"""Compare U-Net outputs of different builds of the library (DAD_LIB_PATH) in subprocesses; also repeatability."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = sys.argv[1] if len(sys.argv) > 1 else "4096"
libs = sys.argv[2:]
child = r'''
import os, sys, torch, numpy as np
sys.path.insert(0, %r)
import bench
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic
w = dict(bench.WORKLOADS["pointmaze"]); B = int(sys.argv[1]); dev = torch.device("cuda", 0); T = 6
w["dim"] = int(os.environ.get("ARCH_DIM", w["dim"])); w["mults"] = tuple(int(v) for v in os.environ.get("ARCH_MULTS", "1,2,4").split(",")); w["H"] = int(os.environ.get("ARCH_H", 32))
net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision="bf16", max_batch=B)
dif = GaussianDiffusion(net, horizon=w["H"], observation_dim=4, action_dim=2, n_timesteps=50)
synthetic.fill_state_dict(dif, 0); dif.to(dev)
x = torch.randn(B, w["H"], T, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
t = torch.full((B,), 7, device=dev, dtype=torch.long)
outs = [net(x, t).cpu().numpy() for _ in range(4)]
print("repeatable:", all(np.array_equal(outs[0], o) for o in outs[1:]))
np.save(sys.argv[2], outs[0])
''' % ROOT
res = {}
for lib in libs:
    env = dict(os.environ)
    spec = lib.split(",")
    env["DAD_TUNING"] = "1"
    env["DAD_LIB_PATH"] = os.path.join(ROOT, "dynamics_aware_diffusion_b200", spec[0])
    for kv in spec[1:]:
        k, v = kv.split("=")
        env[k] = v
    out = tempfile.mktemp(suffix=".npy")
    r = subprocess.run([sys.executable, "-c", child, B, out], env=env, capture_output=True, text=True, timeout=200)
    print(lib, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
    res[lib] = np.load(out)
base = libs[0]
for lib in libs[1:]:
    d = res[lib].astype(np.float64) - res[base]
    per = np.sqrt((d.reshape(d.shape[0], -1) ** 2).sum(1)) / np.sqrt((res[base].reshape(d.shape[0], -1).astype(np.float64) ** 2).sum(1))
    print("%-40s vs %s: rel-L2 %.3e, samples differing %d, max per-sample %.3e" % (lib, base, np.linalg.norm(d) / np.linalg.norm(res[base]), int((per > 0).sum()), per.max()))
