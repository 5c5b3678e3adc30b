"""Host-side (numpy, fp64) construction of the dynamics projector and its fold into one affine map.

  ProjectionMatrixBuilder   P = F pinv(F)            reference: m_diffuser/dynamics/projection.py:11-133
  fit_linear_dynamics       lstsq([X U], X+)          reference: m_diffuser/dynamics/data_driven.py:75-134
  dynamics_residual         mean((tau - tau P)^2)     reference: m_diffuser/losses/__init__.py:161-186
  fold_projection           DynamicsAwarePolicy.apply_projection (guides/policies.py:409-485) as
                            y = x + alpha * (N x + q) on the flattened normalised trajectory
  projection_alphas         _get_projection_alpha for every step (guides/policies.py:358-383)
"""
import numpy as np
import torch


class ProjectionMatrixBuilder:
    """Builds F (trajectory = F [x0, u0..u_{T-1}]) and the orthogonal projector onto its range.
    Concatenated layout [x0, x1, ..., xT, u0, ..., u_{T-1}] as in the reference."""

    def __init__(self, A, B, state_dim: int, action_dim: int, verbose: bool = False):
        A = np.asarray(A, dtype=np.float64)
        B = np.asarray(B, dtype=np.float64)
        assert A.shape == (state_dim, state_dim), f"A shape mismatch: {A.shape}"
        assert B.shape == (state_dim, action_dim), f"B shape mismatch: {B.shape}"
        self.A, self.B, self.state_dim, self.action_dim = A, B, state_dim, action_dim
        self.verbose = verbose

    def _build_F_matrix(self, horizon: int) -> np.ndarray:
        n, m, T = self.state_dim, self.action_dim, horizon
        F = np.zeros(((T + 1) * n + T * m, n + T * m))
        # row block t of the state part is [A^t | A^{t-1}B ... AB B 0 ...]: build it by the recursion
        # row_{t+1} = A row_t, then drop B into the column block of u_t.
        row = np.zeros((n, n + T * m))
        row[:, :n] = np.eye(n)
        F[:n] = row
        for t in range(T):
            row = self.A @ row
            row[:, n + t * m:n + (t + 1) * m] = self.B
            F[(t + 1) * n:(t + 2) * n] = row
        F[(T + 1) * n:, n:] = np.eye(T * m)
        return F

    def get_projection_matrix(self, horizon: int, device=None) -> torch.Tensor:
        """P = F pinv(F) as an fp32 CPU tensor (projection.py:85-120).  `device` (extra, optional): a CUDA device --
        the projector is then built there in fp64 (Gram matrix + Cholesky, `dad_build_projection_matrix`) instead of
        numpy's SVD: the same P to fp32 rounding in milliseconds, for dynamics or horizons that change online."""
        F = self._build_F_matrix(horizon)
        if device is not None:
            from . import _native as N
            dev = torch.device(device)
            if dev.type != "cuda":
                raise ValueError("device must be a CUDA device (the default, None, is the reference's numpy path)")
            F = np.ascontiguousarray(F, dtype=np.float64)
            P32 = np.empty((F.shape[0], F.shape[0]), dtype=np.float32)
            rc = N.lib().dad_build_projection_matrix(dev.index or 0, F.ctypes.data, F.shape[0], F.shape[1], P32.ctypes.data)
            N.check(None, rc)
            return torch.from_numpy(P32)
        P = F @ np.linalg.pinv(F)
        if self.verbose:
            print("projection: F %s, ||P^2-P||_F = %.2e" % (F.shape, np.linalg.norm(P @ P - P)))
        return torch.from_numpy(P).float()

    def verify_projection(self, P: torch.Tensor) -> bool:
        return bool(torch.allclose(P @ P, P, atol=1e-4))


def fit_linear_dynamics(states, actions, next_states, state_dim=None, verbose=False, device=None):
    """Least-squares (A, B) with x+ ~ A x + B u (data_driven.py:107-121).  `device` (extra, optional): a CUDA device --
    the normal equations are then formed and solved there in fp64 (`dad_fit_linear_dynamics`: split-K Gram products +
    Cholesky) instead of numpy's SVD lstsq; the default, None, is the reference's numpy path."""
    states, actions, next_states = (np.ascontiguousarray(a, dtype=np.float64) for a in (states, actions, next_states))
    if state_dim is not None and states.shape[1] > state_dim:
        states, next_states = (np.ascontiguousarray(states[:, :state_dim]),
                               np.ascontiguousarray(next_states[:, :state_dim]))
    n, m = states.shape[1], actions.shape[1]
    if device is not None:
        from . import _native as N
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device (the default, None, is the reference's numpy path)")
        A = np.empty((n, n), dtype=np.float64)
        B = np.empty((n, m), dtype=np.float64)
        rc = N.lib().dad_fit_linear_dynamics(dev.index or 0, states.ctypes.data, actions.ctypes.data,
                                             next_states.ctypes.data, states.shape[0], n, m, A.ctypes.data, B.ctypes.data)
        N.check(None, rc)
        theta = np.vstack([A.T, B.T])
    else:
        theta, *_ = np.linalg.lstsq(np.hstack([states, actions]), next_states, rcond=None)
    if verbose:
        resid = next_states - np.hstack([states, actions]) @ theta
        print("fit_linear_dynamics: R^2 = %.4f" % (1 - (resid ** 2).sum() / ((next_states - next_states.mean(0)) ** 2).sum()))
    return theta[:n].T.copy(), theta[n:].T.copy()


def fold_projection(P, obs_mean, obs_std, action_mean, action_std, state_dim, action_dim, horizon):
    """(Nmat, q) in fp64 such that apply_projection(x, alpha) == x + alpha * (Nmat @ x + q) for the
    row-major flattening x[h*T + j] of a normalised (H, T) trajectory, T = state_dim + action_dim.

    The reference computes, per sample (policies.py:431-485):
        c = U(x)            unnormalise, append x_H := x_{H-1}, concatenate      (affine: c = S x + s0)
        c <- alpha (c P) + (1 - alpha) c
        y = R(c)            drop the appended state, renormalise                 (affine: y = R c - r0, R S = I)
    so y = x + alpha (R P^T S - I) x + alpha R (P^T s0 - s0).
    """
    n, m, H = state_dim, action_dim, horizon
    T = n + m
    P = np.asarray(P, dtype=np.float64)
    Dc = (H + 1) * n + H * m
    if P.shape != (Dc, Dc):
        raise ValueError("projection matrix is %s, expected (%d, %d) for horizon %d" % (P.shape, Dc, Dc, H))
    om, os_ = np.asarray(obs_mean, np.float64).reshape(-1), np.asarray(obs_std, np.float64).reshape(-1)
    am, as_ = np.asarray(action_mean, np.float64).reshape(-1), np.asarray(action_std, np.float64).reshape(-1)
    if om.shape[0] != n or os_.shape[0] != n:
        raise ValueError("the reference projection only runs when observation_dim == state_dim (SURVEY.md F4)")
    S = np.zeros((Dc, H * T))
    s0 = np.zeros(Dc)
    R = np.zeros((H * T, Dc))
    r_shift = np.zeros(H * T)
    for h in range(H + 1):
        src = min(h, H - 1)
        for j in range(n):
            S[h * n + j, src * T + j] = os_[j]
            s0[h * n + j] = om[j]
    for h in range(H):
        for j in range(n):
            R[h * T + j, h * n + j] = 1.0 / os_[j]
            r_shift[h * T + j] = om[j] / os_[j]
        for j in range(m):
            S[(H + 1) * n + h * m + j, h * T + n + j] = as_[j]
            s0[(H + 1) * n + h * m + j] = am[j]
            R[h * T + n + j, (H + 1) * n + h * m + j] = 1.0 / as_[j]
            r_shift[h * T + n + j] = am[j] / as_[j]
    Pt = P.T
    M = R @ Pt @ S
    q = R @ (Pt @ s0) - r_shift
    return M - np.eye(H * T), q


def dynamics_residual(x, P, obs_mean, obs_std, action_mean, action_std, state_dim, action_dim):
    """mean((tau - tau P)^2) in physical space for normalised trajectories x (B, H, T): the reference's dynamics
    violation metric, ProjectionLoss.compute (m_diffuser/losses/__init__.py:161-186).
    A CUDA tensor x is reduced on its device by one fused kernel (`dad_dynamics_residual`: tau is built from x on the
    fly, fp32 products like the reference's torch matmul, fp64 sum); anything else goes through numpy in fp64."""
    if torch.is_tensor(x) and x.is_cuda:
        import ctypes
        from . import _native as N
        xc = x.detach().to(torch.float32).contiguous()
        Pd = torch.as_tensor(P, dtype=torch.float32).to(xc.device).contiguous()
        Bsz, H, T = xc.shape
        if T != state_dim + action_dim or Pd.shape != ((H + 1) * state_dim + H * action_dim,) * 2:
            raise ValueError("x is %s and P is %s: inconsistent with state_dim=%d, action_dim=%d"
                             % (tuple(xc.shape), tuple(Pd.shape), state_dim, action_dim))
        stats = [np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1)) for a in (obs_mean, obs_std, action_mean, action_std)]
        out = ctypes.c_double()
        with torch.cuda.device(xc.device):
            rc = N.lib().dad_dynamics_residual(xc.device.index or 0, xc.data_ptr(), Bsz, H, state_dim, action_dim,
                                               Pd.data_ptr(), *[a.ctypes.data for a in stats], ctypes.byref(out),
                                               torch.cuda.current_stream().cuda_stream)
        N.check(None, rc)
        return float(out.value)
    x = np.asarray(x.detach().cpu() if torch.is_tensor(x) else x, dtype=np.float64)
    B = x.shape[0]
    s = x[:, :, :state_dim] * np.asarray(obs_std, np.float64) + np.asarray(obs_mean, np.float64)
    a = x[:, :, state_dim:state_dim + action_dim] * np.asarray(action_std, np.float64) + np.asarray(action_mean, np.float64)
    s = np.concatenate([s, s[:, -1:, :]], axis=1)                       # duplicated last state (losses/__init__.py:153)
    c = np.concatenate([s.reshape(B, -1), a.reshape(B, -1)], axis=1)
    return float(np.mean((c - c @ np.asarray(P.detach().cpu() if torch.is_tensor(P) else P, dtype=np.float64)) ** 2))


def projection_alphas(n_table, n_timesteps, schedule, strength, betas=None):
    """alpha_i for i in [0, n_table): policies.py:358-383 (progress = i / n_timesteps)."""
    i = np.arange(n_table, dtype=np.float64)
    progress = i / float(n_timesteps)
    if schedule == "constant":
        a = np.full(n_table, float(strength))
    elif schedule == "linear":
        a = strength * (1 - progress)
    elif schedule == "quadratic":
        a = strength * (1 - progress) ** 2
    elif schedule == "noise_schedule":
        b = torch.as_tensor(betas, dtype=torch.float32).cpu()
        a = torch.sqrt(1 - b).double().numpy()[:n_table] * strength     # fp32 sqrt like the reference, then .item()
    else:
        raise ValueError(f"Unknown projection schedule: {schedule}")
    return np.where(a > 0, a, 0.0)      # alpha <= 0: the reference returns x unchanged (policies.py:428-429)
