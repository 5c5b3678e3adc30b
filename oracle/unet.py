"""numpy restatement of TemporalUnet.forward (reference: m_diffuser/models/temporal_unet.py).

Test infrastructure only (see oracle/__init__.py).  Activations are kept
channels-last (B, L, C) so that every convolution is a handful of BLAS matmuls;
weights are consumed in the reference's state_dict layout.
"""
import math

import numpy as np


def mish(x):
    """nn.Mish: x * tanh(softplus(x))  (temporal_unet.py:72,98,158)."""
    sp = np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))
    return x * np.tanh(sp)


def linear(x, w, b):
    """nn.Linear: x @ w.T + b."""
    return x @ w.T + b


def conv1d(x, w, b, stride=1, padding=0):
    """nn.Conv1d on channels-last input.

    x: (B, L, Cin); w: (Cout, Cin, k) (torch layout); returns (B, Lout, Cout).
    Used for Conv1dBlock (k=5,p=2; temporal_unet.py:70), Downsample1d (k=3,s=2,p=1; :40),
    residual / head 1x1 convs (:103,:196).
    """
    B, L, Cin = x.shape
    Cout, Cin_w, k = w.shape
    assert Cin == Cin_w
    Lout = (L + 2 * padding - k) // stride + 1
    xp = np.zeros((B, L + 2 * padding, Cin), dtype=x.dtype)
    xp[:, padding:padding + L] = x
    out = np.zeros((B, Lout, Cout), dtype=x.dtype)
    for tap in range(k):
        xs = xp[:, tap:tap + stride * (Lout - 1) + 1:stride]
        out += (xs.reshape(B * Lout, Cin) @ w[:, :, tap].T).reshape(B, Lout, Cout)
    return out + b


def conv_transpose1d(x, w, b, stride=2, padding=1):
    """nn.ConvTranspose1d on channels-last input (Upsample1d, temporal_unet.py:51).

    x: (B, L, Cin); w: (Cin, Cout, k) (torch layout).  out[o] += x[i] @ w[:,:,kk] with
    o = i*stride - padding + kk.
    """
    B, L, Cin = x.shape
    Cin_w, Cout, k = w.shape
    assert Cin == Cin_w
    Lout = (L - 1) * stride - 2 * padding + k
    out = np.zeros((B, Lout, Cout), dtype=x.dtype)
    for kk in range(k):
        y = (x.reshape(B * L, Cin) @ w[:, :, kk]).reshape(B, L, Cout)
        for i in range(L):
            o = i * stride - padding + kk
            if 0 <= o < Lout:
                out[:, o] += y[:, i]
    return out + b


def group_norm(x, gamma, beta, n_groups=8, eps=1e-5):
    """nn.GroupNorm(8, C) on channels-last input: statistics over (L, C/8) per sample and group
    (biased variance), temporal_unet.py:71."""
    B, L, C = x.shape
    g = x.reshape(B, L, n_groups, C // n_groups)
    mean = g.mean(axis=(1, 3), keepdims=True)
    var = g.var(axis=(1, 3), keepdims=True)
    y = (g - mean) / np.sqrt(var + eps)
    return y.reshape(B, L, C) * gamma + beta


def sinusoidal_pos_emb(t, dim, dtype):
    """SinusoidalPosEmb.forward (temporal_unet.py:19-32); note the (half_dim - 1) divisor."""
    half = dim // 2
    scale = math.log(10000) / (half - 1)
    freq = np.exp(np.arange(half, dtype=dtype) * -scale)
    arg = t.astype(dtype)[:, None] * freq[None, :]
    return np.concatenate([np.sin(arg), np.cos(arg)], axis=-1)


class UnetOracle:
    """Forward pass driven purely by a state_dict (name -> ndarray).

    The architecture (levels, widths, kernel size) is read back from the tensor
    names and shapes, mirroring how the reference builds it (temporal_unet.py:135-197).
    """

    def __init__(self, state_dict, prefix="", dtype=np.float64):
        self.dtype = dtype
        self.w = {k[len(prefix):]: np.asarray(v).astype(dtype) for k, v in state_dict.items()
                  if k.startswith(prefix)}
        self.n_levels = 1 + max(int(k.split(".")[1]) for k in self.w if k.startswith("downs."))
        self.dim = self.w["time_mlp.1.weight"].shape[1]
        self.ksize = self.w["downs.0.0.blocks.0.block.0.weight"].shape[2]

    # -- blocks ---------------------------------------------------------------------------
    def _conv_block(self, x, name):
        """Conv1dBlock: Conv1d(k, pad k//2) -> GroupNorm(8) -> Mish (temporal_unet.py:69-73)."""
        w = self.w
        y = conv1d(x, w[name + ".block.0.weight"], w[name + ".block.0.bias"], padding=self.ksize // 2)
        y = group_norm(y, w[name + ".block.1.weight"], w[name + ".block.1.bias"])
        return mish(y)

    def _res_block(self, x, temb, name):
        """ResidualTemporalBlock.forward (temporal_unet.py:106-122)."""
        w = self.w
        out = self._conv_block(x, name + ".blocks.0")
        tb = linear(mish(temb), w[name + ".time_mlp.1.weight"], w[name + ".time_mlp.1.bias"])
        out = out + tb[:, None, :]
        out = self._conv_block(out, name + ".blocks.1")
        if name + ".residual_conv.weight" in w:
            res = conv1d(x, w[name + ".residual_conv.weight"], w[name + ".residual_conv.bias"])
        else:
            res = x
        return out + res

    def time_embedding(self, t):
        """Global time MLP: SinusoidalPosEmb -> Linear -> Mish -> Linear (temporal_unet.py:155-160)."""
        w = self.w
        e = sinusoidal_pos_emb(np.asarray(t), self.dim, self.dtype)
        e = mish(linear(e, w["time_mlp.1.weight"], w["time_mlp.1.bias"]))
        return linear(e, w["time_mlp.3.weight"], w["time_mlp.3.bias"])

    # -- forward --------------------------------------------------------------------------
    def forward(self, x, t):
        """x: (B, H, T), t: (B,) ints -> (B, H, T)   (temporal_unet.py:199-241).

        The reference transposes to (B, T, H) first (:211); channels-last needs no transpose.
        """
        w = self.w
        x = np.asarray(x).astype(self.dtype)
        temb = self.time_embedding(t)
        skips = []
        for lvl in range(self.n_levels):
            x = self._res_block(x, temb, "downs.%d.0" % lvl)
            x = self._res_block(x, temb, "downs.%d.1" % lvl)
            skips.append(x)
            key = "downs.%d.2.conv.weight" % lvl
            if key in w:          # Identity on the last level (:174)
                x = conv1d(x, w[key], w["downs.%d.2.conv.bias" % lvl], stride=2, padding=1)
        x = self._res_block(x, temb, "mid_block1")
        x = self._res_block(x, temb, "mid_block2")
        for lvl in range(self.n_levels - 1):
            x = np.concatenate([x, skips.pop()], axis=-1)       # cat on channels (:230)
            x = self._res_block(x, temb, "ups.%d.0" % lvl)
            x = self._res_block(x, temb, "ups.%d.1" % lvl)
            # every decoder level upsamples (is_last is never true, :185)
            x = conv_transpose1d(x, w["ups.%d.2.conv.weight" % lvl], w["ups.%d.2.conv.bias" % lvl])
        x = self._conv_block(x, "final_conv.0")
        return conv1d(x, w["final_conv.1.weight"], w["final_conv.1.bias"])
