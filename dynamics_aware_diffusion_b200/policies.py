"""Sampling policies with the reference's classes and signatures (m_diffuser/guides/policies.py), running
the reverse process in the native library.

  GuidedPolicy         conditioning (inpainting) + optional guidance      policies.py:13-223
  MPCPolicy            GuidedPolicy + action_horizon                      policies.py:226-240
  ValueGuidedPolicy    guide_fn from a value nn.Module                    policies.py:243-271
  DynamicsAwarePolicy  per-step projection onto the dynamics-feasible set policies.py:274-485

Without a guide function the whole loop is one `dad_sample` call (CUDA-graph replays).  With one, the
gradient of the user's PyTorch module is taken by autograd each step (as the reference does) and fed to
the fused step kernel.

DynamicsAwarePolicy.sample_loop applies the projection every step (README.md:24-25, 271-275:
x_{i-1} = project(denoise(x_i))).  In the reference `apply_projection` exists but is never called
(SURVEY.md F3); set `policy.project_in_loop = False` to reproduce that behaviour bit-for-bit in spirit.
"""
from typing import Callable, Dict, Optional

import os
import warnings

import numpy as np
import torch
import torch.nn as nn

from . import _native as N
from .diffusion import _run_loop, uniform_timestep
from .projection import fold_projection, projection_alphas


class _CaptureFailed(RuntimeError):
    """The guidance function could not be captured into a CUDA graph."""


class GuidedPolicy(nn.Module):
    def __init__(self, diffusion_model, normalizer, guide_fn: Optional[Callable] = None, guide_weight: float = 1.0,
                 action_horizon: Optional[int] = None):
        super().__init__()
        self.diffusion = diffusion_model
        self.normalizer = normalizer
        self.guide_fn = guide_fn
        self.guide_weight = guide_weight
        self.horizon = diffusion_model.horizon
        self.observation_dim = diffusion_model.observation_dim
        self.action_dim = diffusion_model.action_dim
        self.transition_dim = diffusion_model.transition_dim
        self.action_horizon = action_horizon if action_horizon is not None else 1
        self.action_buffer = []
        # guided loops: capture one step (our kernels + the guide's autograd) in a CUDA graph and replay it; falls
        # back to a per-step host loop, for good, the first time the guide_fn turns out not to be capturable
        self.capture_guidance = os.environ.get("DAD_GUIDED_GRAPH", "1") != "0"
        self._guided_graphs = {}
        self._capture_error = None
        self._pin_cond = self._pin_plan = None

    # ---- hooks overridden by DynamicsAwarePolicy --------------------------------------------------
    def _loop_flags(self, engine):
        return 0

    def _engine(self, device=None):
        return self.diffusion.engine(self.horizon, device)

    # ---- reference API ----------------------------------------------------------------------------
    def apply_conditions(self, x: torch.Tensor, conditions: Dict[int, torch.Tensor]) -> torch.Tensor:
        """x[:, h] = value for every (h, value); writes the whole transition, in place (policies.py:48-63)."""
        for h, val in conditions.items():
            x[:, h] = val
        return x

    def _guidance_grad(self, x, t):
        if self.guide_fn is None or not self.guide_weight > 0:
            return None
        x_req = x.detach().requires_grad_(True)
        with torch.enable_grad():
            score = self.guide_fn(x_req, t)
            (grad,) = torch.autograd.grad(score.sum(), x_req)
        return grad.contiguous()

    @torch.no_grad()
    def p_sample_with_guidance(self, x, t, conditions=None):
        """One guided, conditioned reverse step (policies.py:65-112)."""
        step = uniform_timestep(t)
        eng = self._engine(x.device)
        xc = x.contiguous().float()
        eps = self.diffusion.eps_engine(eng, step).unet_forward(xc, step=step)
        grad = self._guidance_grad(xc, t)
        noise = torch.randn_like(xc)
        flags = 0
        if conditions:
            eng.set_conditions(conditions, xc.shape[0])
            flags |= N.FLAG_CONDITIONS
        out = xc.clone()
        return eng.step(out, eps, step, noise=noise, grad=grad, guide_w=float(self.guide_weight), flags=flags)

    @torch.no_grad()
    def sample_loop(self, batch_size: int = 1, conditions=None, verbose: bool = False,
                    noise: Optional[torch.Tensor] = None, rng: str = "philox", seed: Optional[int] = None,
                    return_trace: bool = False, sample_offset: int = 0):
        """Full conditioned sampling loop (policies.py:114-149); extra keyword arguments as in
        GaussianDiffusion.p_sample_loop."""
        self.diffusion._check_steps()
        device = self.diffusion.betas.device
        x = torch.randn((batch_size, self.horizon, self.transition_dim), device=device)
        eng = self._engine(device)
        flags = self._loop_flags(eng)
        if conditions:
            eng.set_conditions(conditions, batch_size)
            flags |= N.FLAG_CONDITIONS
        guided = self.guide_fn is not None and self.guide_weight > 0
        if not guided:
            return _run_loop(self.diffusion, x, noise, rng, seed, flags, return_trace, sample_offset)
        # guided: autograd supplies the gradient each step, the rest stays native
        S = self.diffusion.n_timesteps
        if conditions:
            x = self.apply_conditions(x, {h: torch.as_tensor(v, device=device) for h, v in conditions.items()})
        x = x.contiguous()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if self.capture_guidance and batch_size <= self.diffusion.model.max_batch:
            try:
                return self._guided_loop_graph(eng, x, noise, rng, seed, flags, return_trace, sample_offset)
            except _CaptureFailed as exc:
                # a guide_fn that cannot run under stream capture (host syncs, CPU work): per-step host loop
                self.capture_guidance = False
                self._capture_error = str(exc)
        trace = []
        for k, i in enumerate(reversed(range(S))):
            t = torch.full((batch_size,), i, device=device, dtype=torch.long)
            eps = self.diffusion.eps_engine(eng, i).unet_forward(x, step=i)
            grad = self._guidance_grad(x, t)
            if noise is not None:
                z = noise[k].to(device, torch.float32).contiguous()
            elif rng == "torch":
                z = torch.randn_like(x)
            else:
                z = None
            eng.step(x, eps, i, noise=z, grad=grad, guide_w=float(self.guide_weight), flags=flags, seed=seed,
                     sample_offset=sample_offset)
            if return_trace:
                trace.append(x.clone())
        return (x, torch.stack(trace)) if return_trace else x

    def _guided_loop_graph(self, eng, x_init, noise, rng, seed, flags, return_trace, sample_offset):
        """The guided loop with ONE captured step replayed S times: our kernels and the value
        model's forward+backward (policies.py:87-97, gradient at x_t) live in the same CUDA graph, the step index
        on the device.  The captured step is cached per (batch, flags, noise mode) and re-captured when the
        native handle reallocates what it points to."""
        S = self.diffusion.n_timesteps
        B = x_init.shape[0]
        device = x_init.device
        mode = "seq" if noise is not None else ("torch" if rng == "torch" else "philox")
        key = (B, int(flags), mode, device.index or 0, id(self.guide_fn))      # the captured step calls THIS guide_fn
        ent = self._guided_graphs.get(key)
        if ent is None or ent["epoch"] != eng.graph_epoch() or ent["engine"] is not eng:
            ent = dict(engine=eng, x=torch.empty_like(x_init), grad=torch.zeros_like(x_init),
                       t=torch.zeros(B, device=device, dtype=torch.long),
                       z=torch.empty_like(x_init) if mode == "torch" else None)
            ent["x"].copy_(x_init)
            ent["t"].fill_(S)
            # warm-up on a side stream (lazy initialisation inside the user's model must not happen under capture)
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                self._guidance_grad(ent["x"], ent["t"] - 1)
                # second, checked run: a guide that synchronises with the host (.item(), .cpu(), CPU tensors) cannot
                # be captured, and a capture that fails half-way leaves torch's CUDA generator unusable -- so find
                # out here, in eager mode, where the only consequence is an exception
                prev = torch.cuda.get_sync_debug_mode()
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")          # "prototype feature" notice
                    torch.cuda.set_sync_debug_mode("error")
                try:
                    self._guidance_grad(ent["x"], ent["t"] - 1)
                except RuntimeError as exc:
                    raise _CaptureFailed("guide_fn synchronises with the host: %s" % str(exc).splitlines()[0]) from exc
                finally:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        torch.cuda.set_sync_debug_mode(prev)
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    ent["t"].sub_(1)
                    eng.loop_unet(B)
                    ent["grad"].copy_(self._guidance_grad(ent["x"], ent["t"]))
                    if ent["z"] is not None:
                        ent["z"].normal_()
                    eng.loop_step(B, flags)
            except Exception as exc:          # torch reports capture violations as RuntimeError
                raise RuntimeError("capturing the guided step into a CUDA graph failed (%s); torch's CUDA RNG state may "
                                   "be unusable in this process now.  Set policy.capture_guidance = False (or "
                                   "DAD_GUIDED_GRAPH=0) for this guide_fn." % str(exc).splitlines()[0]) from exc
            ent["graph"] = graph
            ent["epoch"] = eng.graph_epoch()
            self._guided_graphs = {key: ent}       # one captured step alive at a time
        ent["x"].copy_(x_init)
        trace = torch.empty((S,) + tuple(x_init.shape), device=device) if return_trace else None
        if mode == "seq":
            zseq = noise.to(device, torch.float32).contiguous()
            if zseq.shape[0] < S or tuple(zseq.shape[1:]) != tuple(x_init.shape):
                raise ValueError("noise must be (n_timesteps, B, H, T)")
        else:
            zseq = ent["z"]
        # ill-conditioned leading steps of a bf16 model (GaussianDiffusion.fp32_ill_conditioned_steps): per-step host
        # path with eps from the fp32 sibling; the captured step replays the rest.  Slots are indexed by the step, so
        # the remaining S - lead steps see noise[lead:], trace[lead:] and the same Philox draws.
        lead = 0
        while lead < S and self.diffusion.eps_engine(eng, S - 1 - lead) is not eng:
            i = S - 1 - lead
            ent["t"].fill_(i)
            eps = self.diffusion.eps_engine(eng, i).unet_forward(ent["x"], step=i)
            if mode == "seq":
                z = zseq[lead]
            elif mode == "torch":
                z = ent["z"].normal_()
            else:
                z = None
            eng.step(ent["x"], eps, i, noise=z, grad=self._guidance_grad(ent["x"], ent["t"]),
                     guide_w=float(self.guide_weight), flags=flags, seed=seed, sample_offset=sample_offset)
            if trace is not None:
                trace[lead].copy_(ent["x"])
            lead += 1
        rest = S - lead
        if rest > 0:
            ent["t"].fill_(rest)
            eng.loop_begin(ent["x"], rest, noise=zseq if mode != "seq" else zseq[lead:], noise_single=(mode == "torch"),
                           grad=ent["grad"], guide_w=float(self.guide_weight), flags=flags, seed=seed,
                           sample_offset=sample_offset, trace=None if trace is None else trace[lead:])
            for _ in range(rest):
                ent["graph"].replay()
            eng.loop_replayed(rest)
        out = ent["x"].clone()
        return (out, trace) if return_trace else out

    # ---- environment-facing glue (host side) ----------------------------------------------------------
    def _process_observation(self, observation):
        """dict / sequence observation -> (1, obs_dim) ndarray (policies.py:151-179)."""
        if isinstance(observation, dict):
            keys = observation.keys()
            if "observation" in keys and "desired_goal" in keys:
                state, goal = observation["observation"], observation["desired_goal"]
                wants = self.normalizer.obs_mean.shape[0]
                observation = np.concatenate([state, goal]) if wants == len(state) + len(goal) else state
            elif "observation" in keys:
                observation = observation["observation"]
            elif "achieved_goal" in keys:
                observation = observation["achieved_goal"]
            else:
                observation = np.concatenate([np.asarray(v).flatten() for v in observation.values()])
        return np.asarray(observation).reshape(1, -1)

    def _fill_action_buffer(self, trajectory):
        """First action_horizon+1 actions of plan 0, unnormalised (policies.py:181-191)."""
        plan = trajectory[0].detach().cpu().numpy()
        lo, hi = self.observation_dim, self.observation_dim + self.action_dim
        for h in range(min(self.action_horizon + 1, self.horizon)):
            act = self.normalizer.unnormalize_actions(plan[h, lo:hi].reshape(1, -1))
            self.action_buffer.append(np.asarray(act).flatten())

    def get_action(self, observation, **kwargs) -> np.ndarray:
        """Pop a buffered action, replanning when the buffer is empty (policies.py:193-223)."""
        if self.action_buffer:
            return self.action_buffer.pop(0)
        device = self.diffusion.betas.device
        obs = self.normalizer.normalize_observations(self._process_observation(observation))
        # the condition goes up and the plan comes down through PINNED staging buffers (allocated once): no pageable
        # bounce copies on the per-replan path
        if self._pin_cond is None or self._pin_cond.shape[1] != self.transition_dim:
            self._pin_cond = torch.zeros(1, self.transition_dim).pin_memory()
            self._pin_plan = torch.empty(self.horizon, self.transition_dim).pin_memory()
        self._pin_cond.zero_()
        self._pin_cond[0, :self.observation_dim] = torch.as_tensor(np.asarray(obs), dtype=torch.float32).reshape(-1)
        start = self._pin_cond.to(device, non_blocking=True)
        plan = self.sample_loop(batch_size=1, conditions={0: start}, verbose=False)
        self._pin_plan.copy_(plan[0], non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        self._fill_action_buffer(self._pin_plan[None])
        return self.action_buffer.pop(0)


class MPCPolicy(GuidedPolicy):
    """Plan once, execute `action_horizon` actions, replan."""

    def __init__(self, diffusion_model, normalizer, action_horizon: int = 8):
        super().__init__(diffusion_model, normalizer, action_horizon=action_horizon)


class ValueGuidedPolicy(GuidedPolicy):
    """Guidance from a value network over the observation part of the plan."""

    def __init__(self, diffusion_model, normalizer, value_model: nn.Module, guide_weight: float = 1.0,
                 action_horizon: Optional[int] = None):
        obs_dim = diffusion_model.observation_dim

        def guide_fn(x, t):
            return value_model(x[:, :, :obs_dim]).sum(dim=1)

        super().__init__(diffusion_model, normalizer, guide_fn, guide_weight, action_horizon)
        self.value_model = value_model


class DynamicsAwarePolicy(GuidedPolicy):
    """Sampling with a per-step affine projection onto trajectories consistent with x+ = A x + B u."""

    def __init__(self, diffusion_model, projection_matrix: Optional[torch.Tensor] = None, normalizer=None,
                 state_dim: int = 4, observation_dim: int = 4, action_dim: int = 2, horizon: int = 16,
                 projection_schedule: str = "constant", projection_strength: float = 1.0,
                 action_horizon: Optional[int] = None):
        super().__init__(diffusion_model=diffusion_model, normalizer=normalizer, guide_fn=None, guide_weight=0.0,
                         action_horizon=horizon if action_horizon is None else action_horizon)
        self.projection_matrix = projection_matrix
        self.state_dim, self.observation_dim, self.action_dim = state_dim, observation_dim, action_dim
        self.horizon = horizon
        self.projection_schedule = projection_schedule
        self.projection_strength = projection_strength
        self.n_timesteps = diffusion_model.n_timesteps
        self.device = next(diffusion_model.parameters()).device
        self.project_in_loop = True
        self.project_after_inpaint = False
        stats = ("obs_mean", "obs_std", "action_mean", "action_std")
        for name in stats:
            val = None
            if normalizer is not None:
                val = torch.from_numpy(np.asarray(getattr(normalizer, name))).float().to(self.device)
            setattr(self, name, val)
        self._fold = None          # (Nmat, q) fp64, rebuilt when the matrix or the normaliser changes
        self._fold_tag, self._fold_serial = None, 0
        self._P_device = None      # (matrix identity, fp32 copy of P on the sampling device) for dynamics_residual

    def _get_projection_alpha(self, t: int) -> float:
        betas = self.diffusion.betas if self.projection_schedule == "noise_schedule" else None
        n_table = self.diffusion.betas.shape[0]
        return float(projection_alphas(n_table, self.n_timesteps, self.projection_schedule,
                                       self.projection_strength, betas)[t])

    def unnormalize_states(self, s):
        return s if self.obs_mean is None else s * self.obs_std + self.obs_mean

    def unnormalize_actions(self, a):
        return a if self.action_mean is None else a * self.action_std + self.action_mean

    def normalize_states(self, s):
        return s if self.obs_mean is None else (s - self.obs_mean) / self.obs_std

    def normalize_actions(self, a):
        return a if self.action_mean is None else (a - self.action_mean) / self.action_std

    def _active(self):
        return self.projection_matrix is not None and self.normalizer is not None

    def _projector_inputs_tag(self):
        """Identity of everything the folded map (N, q) is made of: the matrix object (and its in-place version) and
        the normaliser statistics by value.  A new `policy.projection_matrix` or normaliser rebuilds the fold."""
        P = self.projection_matrix
        ptag = (id(P), getattr(P, "_version", None), P.data_ptr() if torch.is_tensor(P) else None)
        stats = tuple(np.asarray(getattr(self.normalizer, n), dtype=np.float64).tobytes()
                      for n in ("obs_mean", "obs_std", "action_mean", "action_std"))
        return (ptag, stats, self.state_dim, self.action_dim, self.horizon)

    def _push_projector(self, eng):
        if self.observation_dim != self.state_dim:
            raise RuntimeError("projection needs observation_dim == state_dim (the reference's apply_projection "
                               "fails otherwise, SURVEY.md F4)")
        ftag = self._projector_inputs_tag()
        if self._fold is None or self._fold_tag != ftag:
            P = self.projection_matrix
            P = P.detach().cpu().numpy() if torch.is_tensor(P) else np.asarray(P)
            self._fold = fold_projection(P, self.normalizer.obs_mean, self.normalizer.obs_std,
                                         self.normalizer.action_mean, self.normalizer.action_std, self.state_dim,
                                         self.action_dim, self.horizon)
            self._fold_tag = ftag
            self._fold_serial += 1
        n_table = self.diffusion.betas.shape[0]
        betas = self.diffusion.betas if self.projection_schedule == "noise_schedule" else None
        alphas = projection_alphas(n_table, self.n_timesteps, self.projection_schedule, self.projection_strength, betas)
        # the handle is shared by every policy built on this diffusion model: the "currently loaded" tag lives on the
        # ENGINE (like the schedule's owner tag), so policy A re-pushes after policy B used the same handle
        sig = (id(self), self._fold_serial, self.projection_schedule, float(self.projection_strength),
               int(self.n_timesteps), n_table, alphas.tobytes())
        if getattr(eng, "projector_tag", None) != sig:
            eng.set_projector(self._fold[0], self._fold[1], alphas)
            eng.projector_tag = sig

    def _loop_flags(self, eng):
        if not (self._active() and self.project_in_loop):
            return 0
        self._push_projector(eng)
        return N.FLAG_PROJECT | (N.FLAG_PROJECT_AFTER_INPAINT if self.project_after_inpaint else 0)

    def dynamics_residual(self, x: torch.Tensor) -> float:
        """The reference's dynamics-violation metric of trajectories x (ProjectionLoss.compute,
        losses/__init__.py:161-186) against this policy's projector; one fused kernel when x lives on the GPU."""
        from .projection import dynamics_residual
        if not self._active():
            raise RuntimeError("dynamics_residual needs a projection matrix and a normalizer")
        if self._P_device is None or self._P_device[0] != self._projector_inputs_tag()[0] or self._P_device[1].device != x.device:
            self._P_device = (self._projector_inputs_tag()[0],
                              torch.as_tensor(self.projection_matrix, dtype=torch.float32).to(x.device).contiguous())
        nz = self.normalizer
        return dynamics_residual(x, self._P_device[1], nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std,
                                 self.state_dim, self.action_dim)

    @torch.no_grad()
    def apply_projection(self, x: torch.Tensor, t: int) -> torch.Tensor:
        """x + alpha_t (N x + q): the reference's unnormalise/concat/@P/blend/split/renormalise chain
        (policies.py:409-485) as one native GEMM.  Returns a new tensor."""
        if not self._active():
            return x
        if self._get_projection_alpha(int(t)) <= 0:
            return x
        eng = self._engine(x.device)
        self._push_projector(eng)
        return eng.project(x.contiguous().float().clone(), int(t))
