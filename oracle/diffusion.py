"""numpy restatement of the DDPM reverse process and the guided / conditioned step.

Reference: m_diffuser/models/diffusion.py (schedules :32-48, buffers :96-128, step :159-251)
and m_diffuser/guides/policies.py (apply_conditions :48-63, p_sample_with_guidance :65-112,
sample_loop :114-149).  Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

BUFFER_NAMES = (
    "betas", "alphas", "alphas_cumprod", "alphas_cumprod_prev",
    "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
    "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
    "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2",
)


def cosine_betas(n, s=0.008, dtype=np.float32):
    """cosine_beta_schedule (diffusion.py:32-41), same op order, in `dtype`."""
    f = dtype
    x = np.linspace(0, n, n + 1, dtype=f)
    ac = np.cos(((x / f(n)) + f(s)) / f(1 + s) * f(np.pi) * f(0.5)) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, f(0.0001), f(0.9999)).astype(f)


def linear_betas(n, beta_start=1e-4, beta_end=0.02, dtype=np.float32):
    """linear_beta_schedule (diffusion.py:44-48)."""
    return np.linspace(beta_start, beta_end, n, dtype=dtype)


def make_buffers(n, schedule="cosine", dtype=np.float32):
    """The 12 registered schedule buffers (diffusion.py:96-128)."""
    f = dtype
    if schedule == "cosine":
        betas = cosine_betas(n, dtype=f)
    elif schedule == "linear":
        betas = linear_betas(n, dtype=f)
    else:
        raise ValueError("Unknown beta schedule: %s" % schedule)
    alphas = (f(1.0) - betas).astype(f)
    ac = np.cumprod(alphas, dtype=f)
    ac_prev = np.concatenate([np.ones(1, dtype=f), ac[:-1]])
    pv = betas * (f(1.0) - ac_prev) / (f(1.0) - ac)
    out = {
        "betas": betas, "alphas": alphas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(f(1.0) - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(f(1.0) / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(f(1.0) / ac - f(1)),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": np.log(np.maximum(pv, f(1e-20))),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (f(1.0) - ac),
        "posterior_mean_coef2": (f(1.0) - ac_prev) * np.sqrt(alphas) / (f(1.0) - ac),
    }
    return {k: v.astype(f) for k, v in out.items()}


def apply_conditions(x, conditions):
    """GuidedPolicy.apply_conditions (policies.py:48-63): x[:, h] = val, whole transition, in place."""
    if conditions:
        for h, val in conditions.items():
            x[:, h] = np.asarray(val, dtype=x.dtype)
    return x


class DiffusionOracle:
    """Reverse process over a `UnetOracle` (or any callable eps_fn(x, t_vec))."""

    def __init__(self, buffers, eps_fn, dtype=np.float64, clip_denoised=True, predict_epsilon=True):
        self.b = {k: np.asarray(buffers[k]).astype(dtype) for k in BUFFER_NAMES}
        self.eps_fn = eps_fn
        self.dtype = dtype
        self.clip_denoised = clip_denoised
        self.predict_epsilon = predict_epsilon
        self.n_timesteps = len(self.b["betas"])

    def p_mean_variance(self, x, i, model_out=None):
        """diffusion.py:182-203 at a uniform timestep i.  Returns (mean, logvar, model_out)."""
        b = self.b
        if model_out is None:
            model_out = self.eps_fn(x, np.full((x.shape[0],), i, dtype=np.int64))
        if self.predict_epsilon:
            x0 = b["sqrt_recip_alphas_cumprod"][i] * x - b["sqrt_recipm1_alphas_cumprod"][i] * model_out
        else:
            x0 = model_out
        if self.clip_denoised:
            x0 = np.clip(x0, -1.0, 1.0)
        mean = b["posterior_mean_coef1"][i] * x0 + b["posterior_mean_coef2"][i] * x
        return mean, b["posterior_log_variance_clipped"][i], model_out

    def p_sample(self, x, i, noise, model_out=None):
        """diffusion.py:205-223."""
        mean, logvar, _ = self.p_mean_variance(x, i, model_out)
        mask = 0.0 if i == 0 else 1.0
        return mean + mask * np.exp(0.5 * logvar) * noise

    def guided_step(self, x, i, noise, conditions=None, grad=None, guide_weight=0.0, model_out=None):
        """GuidedPolicy.p_sample_with_guidance (policies.py:65-112).  `grad` is d(sum guide)/dx at x_t
        (the reference gets it from autograd, :87-94); scaled by exp(logvar) = sigma^2 (:97)."""
        mean, logvar, _ = self.p_mean_variance(x, i, model_out)
        if grad is not None and guide_weight > 0:
            mean = mean + guide_weight * np.exp(logvar) * grad
        mask = 0.0 if i == 0 else 1.0
        x_prev = mean + mask * np.exp(0.5 * logvar) * noise
        return apply_conditions(x_prev, conditions)

    def p_sample_loop(self, x_init, noises, trace=None):
        """diffusion.py:225-251 with injected x_S and z_i (noises[k] is used at step i = S-1-k)."""
        x = np.array(x_init, dtype=self.dtype)
        for k, i in enumerate(reversed(range(self.n_timesteps))):
            x = self.p_sample(x, i, np.asarray(noises[k], dtype=self.dtype))
            if trace is not None:
                trace.append(x.copy())
        return x

    def sample_loop(self, x_init, noises, conditions=None, projector=None, grad_fn=None,
                    guide_weight=0.0, project_after_inpaint=False, trace=None):
        """GuidedPolicy.sample_loop (policies.py:114-149); with `projector` (an
        oracle.projection.ProjectionOracle) it is the dynamics-aware composition of SURVEY.md 8(c):
        denoise (no conditions) -> apply_projection(x, i) -> apply_conditions."""
        x = np.array(x_init, dtype=self.dtype)
        x = apply_conditions(x, conditions)
        for k, i in enumerate(reversed(range(self.n_timesteps))):
            z = np.asarray(noises[k], dtype=self.dtype)
            g = grad_fn(x, i) if grad_fn is not None else None
            if projector is None:
                x = self.guided_step(x, i, z, conditions, g, guide_weight)
            elif project_after_inpaint:
                x = self.guided_step(x, i, z, conditions, g, guide_weight)
                x = projector.apply(x, i)
            else:
                x = self.guided_step(x, i, z, None, g, guide_weight)
                x = projector.apply(x, i)
                x = apply_conditions(x, conditions)
            if trace is not None:
                trace.append(x.copy())
        return x
