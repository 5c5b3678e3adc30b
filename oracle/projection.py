"""numpy restatement of the dynamics projector and its per-step application.

Reference: m_diffuser/dynamics/projection.py (F :43-83, P = F F^+ :85-120),
m_diffuser/guides/policies.py (_get_projection_alpha :358-383, apply_projection :409-485),
m_diffuser/dynamics/data_driven.py (fit_linear_dynamics :75-134),
m_diffuser/losses/__init__.py (dynamics residual :161-186).
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np


def build_F(A, B, horizon):
    """tau = [x_0..x_T, u_0..u_{T-1}] = F [x_0, u_0..u_{T-1}]   (projection.py:43-83)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    n, m = B.shape
    T = horizon
    F = np.zeros(((T + 1) * n + T * m, n + T * m))
    Ap = np.eye(n)
    powers_B = []
    for t in range(T + 1):
        F[t * n:(t + 1) * n, :n] = Ap                    # free response A^t
        if t < T:
            powers_B.append(Ap @ B)                      # A^t B
            Ap = Ap @ A
    for t in range(1, T + 1):
        for tau in range(t):
            F[t * n:(t + 1) * n, n + tau * m:n + (tau + 1) * m] = powers_B[t - tau - 1]
    F[(T + 1) * n:, n:] = np.eye(T * m)
    return F


def projection_matrix(A, B, horizon):
    """P = F pinv(F), fp64, then cast to fp32 as the reference does (projection.py:104-120)."""
    F = build_F(A, B, horizon)
    P = F @ np.linalg.pinv(F)
    return P.astype(np.float32)


def fit_linear_dynamics(states, actions, next_states):
    """lstsq([X U], X+) -> A, B   (data_driven.py:107-121)."""
    n = states.shape[1]
    Phi = np.hstack([states, actions])
    Theta = np.linalg.lstsq(Phi, next_states, rcond=None)[0]
    return Theta[:n, :].T, Theta[n:, :].T


def projection_alpha(i, n_timesteps, schedule, strength, betas=None):
    """_get_projection_alpha (policies.py:358-383).  For 'noise_schedule' the reference takes
    sqrt(1 - betas[i]) in fp32 torch and converts with .item() (:377-378)."""
    progress = i / n_timesteps
    if schedule == "constant":
        return strength
    if schedule == "linear":
        return strength * (1 - progress)
    if schedule == "quadratic":
        return strength * (1 - progress) ** 2
    if schedule == "noise_schedule":
        b = np.float32(betas[i])
        return float(np.sqrt(np.float32(1) - b)) * strength
    raise ValueError("Unknown projection schedule: %s" % schedule)


class ProjectionOracle:
    """Literal restatement of DynamicsAwarePolicy.apply_projection (policies.py:409-485):
    unnormalise -> duplicate last state -> concatenate -> @ P -> alpha blend -> split ->
    drop last state -> renormalise.  Requires observation_dim == state_dim."""

    def __init__(self, P, obs_mean, obs_std, action_mean, action_std, state_dim, action_dim,
                 horizon, n_timesteps, schedule="constant", strength=1.0, betas=None,
                 dtype=np.float64):
        self.P = np.asarray(P).astype(dtype)
        self.om, self.os = np.asarray(obs_mean).astype(dtype), np.asarray(obs_std).astype(dtype)
        self.am, self.as_ = np.asarray(action_mean).astype(dtype), np.asarray(action_std).astype(dtype)
        self.n, self.m, self.H = state_dim, action_dim, horizon
        self.S, self.schedule, self.strength, self.betas = n_timesteps, schedule, strength, betas
        self.dtype = dtype

    def alpha(self, i):
        return projection_alpha(i, self.S, self.schedule, self.strength, self.betas)

    def to_concat_physical(self, x):
        B = x.shape[0]
        s = x[:, :, :self.n] * self.os + self.om
        a = x[:, :, self.n:] * self.as_ + self.am
        s_ext = np.concatenate([s, s[:, -1:, :]], axis=1)
        return np.concatenate([s_ext.reshape(B, -1), a.reshape(B, -1)], axis=1)

    def apply(self, x, i):
        a = self.alpha(i)
        if a <= 0:
            return x
        B = x.shape[0]
        c = self.to_concat_physical(np.asarray(x, dtype=self.dtype))
        c = a * (c @ self.P) + (1 - a) * c
        ns = (self.H + 1) * self.n
        s = c[:, :ns].reshape(B, self.H + 1, self.n)[:, :-1, :]
        u = c[:, ns:].reshape(B, self.H, self.m)
        s = (s - self.om) / self.os
        u = (u - self.am) / self.as_
        return np.concatenate([s, u], axis=-1)

    def residual(self, x):
        """ProjectionLoss.compute (losses/__init__.py:161-186): mean((tau - tau P)^2), physical space."""
        c = self.to_concat_physical(np.asarray(x, dtype=self.dtype))
        return float(np.mean((c - c @ self.P) ** 2))
