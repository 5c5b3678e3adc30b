"""TemporalUnet with the reference's constructor, parameters and state_dict layout, executed by the
native sm_100a library.  Mirrors m_diffuser/models/temporal_unet.py (class names and nn.Module nesting are
dictated by the checkpoint key layout); the forward pass does not run PyTorch ops.
"""
import os
from typing import Optional

import torch
import torch.nn as nn

from .engine import create_engine_auto

_DEFAULT_MAX_BATCH = int(os.environ.get("DAD_MAX_BATCH", "4096"))


class SinusoidalPosEmb(nn.Module):
    """Parameter-free placeholder at time_mlp[0]; the table is built on the device (temporal_unet.py:12-32)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim


def _conv_gn(cin, cout, k):
    # children indexed 0 (conv) and 1 (norm); the reference's Mish at index 2 holds no parameters
    return nn.Sequential(nn.Conv1d(cin, cout, k, padding=k // 2), nn.GroupNorm(8, cout), nn.Mish())


class Conv1dBlock(nn.Module):
    def __init__(self, inp_channels, out_channels, kernel_size=3, n_groups=8):
        super().__init__()
        assert n_groups == 8, "the native kernels implement GroupNorm(8, C) as the reference uses it"
        self.block = _conv_gn(inp_channels, out_channels, kernel_size)


class ResidualTemporalBlock(nn.Module):
    def __init__(self, inp_channels, out_channels, embed_dim=128, kernel_size=5):
        super().__init__()
        self.blocks = nn.ModuleList([Conv1dBlock(inp_channels, out_channels, kernel_size),
                                     Conv1dBlock(out_channels, out_channels, kernel_size)])
        self.time_mlp = nn.Sequential(nn.Mish(), nn.Linear(embed_dim, out_channels))
        self.residual_conv = (nn.Conv1d(inp_channels, out_channels, 1)
                              if inp_channels != out_channels else nn.Identity())


class Downsample1d(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv1d(dim, dim, kernel_size=3, stride=2, padding=1)


class Upsample1d(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv = nn.ConvTranspose1d(dim, dim, kernel_size=4, stride=2, padding=1)


class TemporalUnet(nn.Module):
    """Drop-in for m_diffuser.models.temporal_unet.TemporalUnet (constructor: temporal_unet.py:135-140).

    Extra, optional knobs (not in the reference): `precision` in {'auto','bf16','fp32'}, `max_batch`
    (workspace capacity in samples; bigger batches are processed in chunks) and `latency_max_batch` (batches up to
    this size run the latency kernels written for get_action's single plan; None = library default, 0 = never).
    """

    def __init__(self, transition_dim: int, dim: int = 128, dim_mults: tuple = (1, 2, 4, 8), kernel_size: int = 5,
                 time_dim: Optional[int] = None, precision: str = "auto", max_batch: Optional[int] = None,
                 latency_max_batch: Optional[int] = None):
        super().__init__()
        self.transition_dim = transition_dim
        self.dim, self.dim_mults, self.kernel_size = dim, tuple(dim_mults), kernel_size
        self.time_dim = time_dim or dim
        self.precision = precision
        self.max_batch = max_batch or _DEFAULT_MAX_BATCH
        self.latency_max_batch = latency_max_batch
        td = self.time_dim
        # construction order follows the reference so that default init under a seed is identical
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, td * 4), nn.Mish(), nn.Linear(td * 4, td))
        widths = [transition_dim] + [dim * m for m in self.dim_mults]
        pairs = list(zip(widths[:-1], widths[1:]))
        n = len(pairs)
        self.downs = nn.ModuleList()
        for i, (ci, co) in enumerate(pairs):
            self.downs.append(nn.ModuleList([
                ResidualTemporalBlock(ci, co, embed_dim=td, kernel_size=kernel_size),
                ResidualTemporalBlock(co, co, embed_dim=td, kernel_size=kernel_size),
                Downsample1d(co) if i < n - 1 else nn.Identity()]))
        mid = widths[-1]
        self.mid_block1 = ResidualTemporalBlock(mid, mid, embed_dim=td, kernel_size=kernel_size)
        self.mid_block2 = ResidualTemporalBlock(mid, mid, embed_dim=td, kernel_size=kernel_size)
        self.ups = nn.ModuleList()
        for ci, co in reversed(pairs[1:]):
            # every decoder level upsamples: the reference's is_last test never fires (temporal_unet.py:185)
            self.ups.append(nn.ModuleList([
                ResidualTemporalBlock(co * 2, ci, embed_dim=td, kernel_size=kernel_size),
                ResidualTemporalBlock(ci, ci, embed_dim=td, kernel_size=kernel_size),
                Upsample1d(ci)]))
        self.final_conv = nn.Sequential(Conv1dBlock(dim, dim, kernel_size=kernel_size), nn.Conv1d(dim, transition_dim, 1))
        self._engines = {}
        self._n_timesteps = 1000      # time-table length for a stand-alone forward; GaussianDiffusion sets it to its schedule length
        self._force_reload, self._reload_count = False, 0
        self._diffusion_cfg = dict(predict_epsilon=True, clip_denoised=True)

    # ---- native engine management ---------------------------------------------------------------
    def _weights_version(self):
        """Change tag of the parameters: (storage address, autograd version) of every tensor catches optimizer steps,
        `load_state_dict` and `p.data = new`; a device checksum catches in-place writes THROUGH `.data`
        (`p.data.copy_()`, `p.data.mul_()`: EMA.apply_shadow / Trainer.update_ema / NCCL broadcasts), which bump
        neither.  One fused norm kernel + one scalar read per call."""
        params = [p for p in self.parameters()]
        tag = tuple((p.data_ptr(), p._version) for p in params)
        if self._force_reload:
            self._force_reload = False
            self._reload_count += 1
        dev = [p.detach() for p in params if p.is_cuda]
        check = 0.0
        if dev:
            norms = torch.stack(torch._foreach_norm(dev, 1)).double()
            check = float((norms * torch.arange(1, norms.numel() + 1, device=norms.device, dtype=torch.float64)).sum())
        return (tag, check, self._reload_count)

    def invalidate(self):
        """Force the native handles to re-pack the weights on their next use (call after writing parameters in a way
        PyTorch does not track)."""
        self._force_reload = True

    refresh_weights = invalidate

    def engine(self, horizon, device, n_timesteps=None, min_batch=1, precision=None):
        """The native handle for (horizon, device); rebuilt when shapes change, re-packed when weights change.
        `precision` overrides the module's (the fp32 sibling that evaluates ill-conditioned steps of a bf16 model)."""
        n_t = int(n_timesteps or self._n_timesteps)
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("TemporalUnet forward needs CUDA tensors: there is no CPU fallback in this implementation")
        precision = precision or self.precision
        key = (int(horizon), device.index or 0, n_t, precision, tuple(sorted(self._diffusion_cfg.items())))
        ent = self._engines.get(key)
        if ent is None:
            eng = create_engine_auto(precision, transition_dim=self.transition_dim, dim=self.dim,
                                     dim_mults=self.dim_mults, kernel_size=self.kernel_size, time_dim=self.time_dim,
                                     horizon=horizon, n_timesteps=n_t, max_batch=self.max_batch, device=device,
                                     **self._diffusion_cfg)
            if self.latency_max_batch is not None:
                eng.set_latency_batch(self.latency_max_batch)
            ent = {"engine": eng, "version": None}
            self._engines[key] = ent
        ver = self._weights_version()
        if ent["version"] != ver:
            ent["engine"].load_unet_state(self.state_dict().items())
            ent["version"] = ver
            ent["schedule_owner"] = None
        return ent["engine"], ent

    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        """x: (B, H, T) fp32 CUDA, time: (B,) integer timesteps -> (B, H, T)   (temporal_unet.py:199-241)."""
        if x.dim() != 3 or x.shape[2] != self.transition_dim:
            raise ValueError("x must be (batch, horizon, %d)" % self.transition_dim)
        time = time.reshape(-1)
        if time.is_floating_point():
            if not bool((time == time.round()).all()):
                raise ValueError("time must hold integer timesteps: the time-embedding tables are indexed by step")
            time = time.long()
        lo, hi = (int(v) for v in torch.aminmax(time))
        if lo < 0 or hi >= self._n_timesteps:
            raise IndexError("timesteps must lie in [0, %d) (got %d..%d): the tables are built for the diffusion's "
                             "schedule length" % (self._n_timesteps, lo, hi))
        eng, _ = self.engine(x.shape[1], x.device)
        return eng.unet_forward(x.contiguous().float(), t=time)
