"""Synthetic inputs shared by tests, the golden-vector generator and bench.py:
seeded weights, a duck-typed stand-in for the reference's missing DatasetNormalizer, known and
data-driven linear dynamics."""
import numpy as np
import torch


class SyntheticNormalizer:
    """Duck-typed DatasetNormalizer (the reference's class is not in its tree): numpy attrs
    obs_mean/obs_std/action_mean/action_std and the two methods the policies call."""

    def __init__(self, obs_dim, action_dim, seed=7):
        rng = np.random.default_rng(seed)
        self.obs_mean = rng.normal(0, 1, obs_dim).astype(np.float32)
        self.obs_std = rng.uniform(0.5, 1.5, obs_dim).astype(np.float32)
        self.action_mean = (0.1 * rng.normal(0, 1, action_dim)).astype(np.float32)
        self.action_std = rng.uniform(0.5, 1.5, action_dim).astype(np.float32)

    def normalize_observations(self, o):
        return (np.asarray(o, dtype=np.float32) - self.obs_mean) / self.obs_std

    def unnormalize_actions(self, a):
        return np.asarray(a, dtype=np.float32) * self.action_std + self.action_mean


def double_integrator(dt=0.1):
    """The reference's analytical PointMaze dynamics (m_diffuser/dynamics/extractor.py:119-131)."""
    A = np.eye(4)
    A[0, 2] = A[1, 3] = dt
    B = np.zeros((4, 2))
    B[0, 0] = B[1, 1] = 0.5 * dt ** 2
    B[2, 0] = B[3, 1] = dt
    return A, B


def random_linear_system(n, m, seed=11, n_transitions=100_000):
    """Ground-truth (A, B) and noisy transitions for the data-driven configs."""
    rng = np.random.default_rng(seed)
    A = np.eye(n) + 0.05 * rng.normal(0, 1, (n, n)) / np.sqrt(n)
    B = 0.1 * rng.normal(0, 1, (n, m))
    X = rng.normal(0, 1, (n_transitions, n))
    U = rng.normal(0, 1, (n_transitions, m))
    Xn = X @ A.T + U @ B.T + 1e-3 * rng.normal(0, 1, (n_transitions, n))
    return A, B, X, U, Xn


def fill_state_dict(module: torch.nn.Module, seed: int):
    """Deterministic, platform-independent weights for every parameter of `module` (numpy PCG64):
    conv/linear weights ~ U(-a, a) with a = 1/sqrt(fan_in), biases ~ U(-a, a), GroupNorm gamma ~ U(0.5, 1.5),
    beta ~ U(-0.2, 0.2).  Buffers (schedules) are left alone.  Returns the numpy dict that was loaded."""
    rng = np.random.default_rng(seed)
    out = {}
    with torch.no_grad():
        for name, p in module.named_parameters():
            shape = tuple(p.shape)
            if p.dim() >= 2:
                fan_in = int(np.prod(shape[1:]))
                a = 1.0 / np.sqrt(fan_in)
                v = rng.uniform(-a, a, shape)
            elif ".block.1.weight" in name:
                v = rng.uniform(0.5, 1.5, shape)
            elif ".block.1.bias" in name:
                v = rng.uniform(-0.2, 0.2, shape)
            else:
                v = rng.uniform(-0.1, 0.1, shape)
            v = v.astype(np.float32)
            p.copy_(torch.from_numpy(v))
            out[name] = v
    return out
