"""Diagnostic (not a test): per-case eps and per-step errors of both precisions vs the golden vectors,
with stock torch.autocast(bf16) beside them for calibration."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from oracle import torch_port as torch_ref  # noqa: E402
from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for name, c in helpers.CASES.items():
    sd, _ = helpers.make_state_dict(c)
    g = helpers.load_golden(name)
    w = {k[len("model."):]: torch.from_numpy(v).to(dev) for k, v in sd.items() if k.startswith("model.")}
    x = torch.from_numpy(g["x_init"]).to(dev)
    for precision in ("fp32", "bf16"):
        net = TemporalUnet(helpers.case_T(c), dim=c["dim"], dim_mults=c["mults"], precision=precision, max_batch=64)
        dif = GaussianDiffusion(net, horizon=c["H"], observation_dim=c["n"], action_dim=c["m"], n_timesteps=c["S"], beta_schedule=c["beta"])
        dif.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        dif.to(dev)
        errs = []
        for i, want in zip(g["unet_steps"], g["unet_eps"]):
            got = net(x, torch.full((c["B"],), int(i), device=dev, dtype=torch.long))
            errs.append(helpers.rel_l2(got.cpu().numpy(), want))
        steps = []
        for k, i in enumerate(reversed(range(c["S"]))):
            x_in = g["x_init"] if k == 0 else g["trace_plain"][k - 1]
            z = torch.from_numpy(g["noise"][k]).to(dev)
            real = torch.randn_like
            torch.randn_like = lambda t, **kw: z
            try:
                got = dif.p_sample(torch.from_numpy(x_in).to(dev), torch.full((c["B"],), i, device=dev, dtype=torch.long))
            finally:
                torch.randn_like = real
            steps.append(helpers.rel_l2(got.cpu().numpy(), g["trace_plain"][k]))
        print("%-10s %-5s eps %s | step max %.2e first %.2e rest max %.2e" % (
            name, precision, " ".join("%.2e" % e for e in errs), max(steps), steps[0], max(steps[1:])))
    errs = []
    for i, want in zip(g["unet_steps"], g["unet_eps"]):
        t = torch.full((c["B"],), int(i), device=dev, dtype=torch.long)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got = torch_ref.unet_forward(w, x, t).float()
        f32 = torch_ref.unet_forward(w, x, t)
        errs.append((helpers.rel_l2(got.cpu().numpy(), want), helpers.rel_l2(f32.cpu().numpy(), want)))
    print("%-10s torch autocast-bf16 eps %s | torch fp32 eps %s" % (
        name, " ".join("%.2e" % e[0] for e in errs), " ".join("%.2e" % e[1] for e in errs)))
