"""PyTorch (torch.nn.functional, fp32) restatement of the reference's sampling path, driven by a state_dict.

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure PyTorch and cannot travel to
the GPU box (/root/reference does not exist there), so this file restates its forward op for op with the same
ATen calls -- conv1d, conv_transpose1d, group_norm, mish, linear, gather-style coefficient lookups, the
15-op projection chain -- which makes it (a) the timed CPU baseline of bench.py (`cpu_baseline.kind = "port"`
and `--impl reference`): same kernels (oneDNN) and the same op count as the reference's CPU sampler, and
(b) the stock-bf16 calibration (torch.autocast) for the GPU parity tests.  It is pinned against the golden
vectors of the live reference in tests/test_oracle_golden.py.

  unet_forward            TemporalUnet.forward                        m_diffuser/models/temporal_unet.py:199-241
  p_mean_variance/p_step  GaussianDiffusion.p_mean_variance/p_sample  m_diffuser/models/diffusion.py:159-223
                          + GuidedPolicy.p_sample_with_guidance       m_diffuser/guides/policies.py:65-112
  apply_projection        DynamicsAwarePolicy.apply_projection        m_diffuser/guides/policies.py:409-485
  sample_loop             GuidedPolicy.sample_loop + the dynamics-aware composition
"""
import math

import torch
import torch.nn.functional as F


def _block(x, w, name, k):
    y = F.conv1d(x, w[name + ".block.0.weight"], w[name + ".block.0.bias"], padding=k // 2)
    y = F.group_norm(y, 8, w[name + ".block.1.weight"], w[name + ".block.1.bias"])
    return F.mish(y)


def _res(x, temb, w, name, k):
    out = _block(x, w, name + ".blocks.0", k)
    out = out + F.linear(F.mish(temb), w[name + ".time_mlp.1.weight"], w[name + ".time_mlp.1.bias"])[:, :, None]
    out = _block(out, w, name + ".blocks.1", k)
    if name + ".residual_conv.weight" in w:
        x = F.conv1d(x, w[name + ".residual_conv.weight"], w[name + ".residual_conv.bias"])
    return out + x


def unet_forward(w, x, t):
    """w: {key relative to TemporalUnet: tensor}; x (B,H,T); t (B,) long."""
    dim = w["time_mlp.1.weight"].shape[1]
    k = w["downs.0.0.blocks.0.block.0.weight"].shape[2]
    n_levels = 1 + max(int(s.split(".")[1]) for s in w if s.startswith("downs."))
    half = dim // 2
    freq = torch.exp(torch.arange(half, device=x.device) * -(math.log(10000) / (half - 1)))
    e = t[:, None].float() * freq[None]
    e = torch.cat([e.sin(), e.cos()], dim=-1)
    temb = F.linear(F.mish(F.linear(e, w["time_mlp.1.weight"], w["time_mlp.1.bias"])), w["time_mlp.3.weight"], w["time_mlp.3.bias"])
    x = x.transpose(1, 2)
    skips = []
    for l in range(n_levels):
        x = _res(x, temb, w, "downs.%d.0" % l, k)
        x = _res(x, temb, w, "downs.%d.1" % l, k)
        skips.append(x)
        if "downs.%d.2.conv.weight" % l in w:
            x = F.conv1d(x, w["downs.%d.2.conv.weight" % l], w["downs.%d.2.conv.bias" % l], stride=2, padding=1)
    x = _res(x, temb, w, "mid_block1", k)
    x = _res(x, temb, w, "mid_block2", k)
    for l in range(n_levels - 1):
        x = torch.cat([x, skips.pop()], dim=1)
        x = _res(x, temb, w, "ups.%d.0" % l, k)
        x = _res(x, temb, w, "ups.%d.1" % l, k)
        x = F.conv_transpose1d(x, w["ups.%d.2.conv.weight" % l], w["ups.%d.2.conv.bias" % l], stride=2, padding=1)
    x = _block(x, w, "final_conv.0", k)
    x = F.conv1d(x, w["final_conv.1.weight"], w["final_conv.1.bias"])
    return x.transpose(1, 2)


def p_step(sd, w, x, i, noise, conditions=None, grad=None, guide_weight=0.0):
    """One reverse step at uniform timestep i (diffusion.py:182-223, policies.py:84-110)."""
    B = x.shape[0]
    t = torch.full((B,), i, device=x.device, dtype=torch.long)

    def ext(name):
        return sd[name].gather(-1, t).reshape(B, 1, 1)          # extract(), diffusion.py:15-29

    eps = unet_forward(w, x, t)
    x0 = ext("sqrt_recip_alphas_cumprod") * x - ext("sqrt_recipm1_alphas_cumprod") * eps
    x0 = torch.clamp(x0, -1.0, 1.0)
    mean = ext("posterior_mean_coef1") * x0 + ext("posterior_mean_coef2") * x
    logvar = ext("posterior_log_variance_clipped")
    if grad is not None and guide_weight > 0:
        mean = mean + guide_weight * logvar.exp() * grad
    mask = (t != 0).float().view(-1, 1, 1)
    out = mean + mask * torch.exp(0.5 * logvar) * noise
    if conditions:
        for h, val in conditions.items():
            out[:, h] = val
    return out


def apply_projection(x, P, alpha, nz, n, m, H):
    """The reference's literal chain (policies.py:431-485); nz = (obs_mean, obs_std, act_mean, act_std) tensors."""
    if alpha <= 0:
        return x
    B = x.shape[0]
    om, os_, am, as_ = nz
    s = x[:, :, :n] * os_ + om
    a = x[:, :, n:] * as_ + am
    s = torch.cat([s, s[:, -1:, :]], dim=1)
    c = torch.cat([s.reshape(B, -1), a.reshape(B, -1)], dim=1)
    c = alpha * (c @ P) + (1 - alpha) * c
    ns = (H + 1) * n
    s = c[:, :ns].reshape(B, H + 1, n)[:, :-1, :]
    a = c[:, ns:].reshape(B, H, m)
    return torch.cat([(s - om) / os_, (a - am) / as_], dim=-1)


@torch.no_grad()
def sample_loop(sd, x, noises, conditions=None, projector=None, steps=None):
    """x_S -> x_0 (or the first `steps` iterations).  projector = dict(P, alphas, nz, n, m, H) or None.
    noises: callable k -> z, or a sequence."""
    w = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    S = sd["betas"].shape[0]
    if conditions:
        for h, val in conditions.items():
            x[:, h] = val
    for k, i in enumerate(reversed(range(S))):
        if steps is not None and k >= steps:
            break
        z = noises(k) if callable(noises) else noises[k]
        if projector is None:
            x = p_step(sd, w, x, i, z, conditions)
        else:
            x = p_step(sd, w, x, i, z, None)
            x = apply_projection(x, projector["P"], projector["alphas"][i], projector["nz"], projector["n"],
                                 projector["m"], projector["H"])
            if conditions:
                for h, val in conditions.items():
                    x[:, h] = val
    return x
