"""Generate tests/golden/*.npz by running the UNMODIFIED reference (m_diffuser, PyTorch CPU fp32).

Run once in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The reference has no golden vectors of its own for this path, so parity is pinned to
outputs of the reference itself: its TemporalUnet / GaussianDiffusion / GuidedPolicy /
ValueGuidedPolicy / DynamicsAwarePolicy / ProjectionMatrixBuilder / fit_linear_dynamics are
imported through oracle/ref_shim.py, loaded with the deterministic weights of tests/helpers.py, and fed
pre-generated noise by temporarily replacing torch.randn / torch.randn_like (the reference looks both
up through the torch module at call time: diffusion.py:218,241, policies.py:100,134).

Every file stores inputs, per-step traces x_i (S,B,H,T), a few raw U-Net outputs and, for the
dynamics-aware composition of SURVEY.md 8(c), the per-step dynamics residual computed as
ProjectionLoss does (losses/__init__.py:161-186).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from oracle import ref_shim  # noqa: E402


@contextlib.contextmanager
def injected_noise(x_init, zs):
    """torch.randn -> x_init once; torch.randn_like -> successive zs."""
    it = iter([torch.from_numpy(np.array(z)) for z in zs])
    real_randn, real_like = torch.randn, torch.randn_like
    torch.randn = lambda *a, **k: torch.from_numpy(np.array(x_init))
    torch.randn_like = lambda t, **k: next(it)
    try:
        yield
    finally:
        torch.randn, torch.randn_like = real_randn, real_like


class ValueModel(torch.nn.Module):
    """V(obs) = tanh(obs . w): (B, H, n) -> (B, H); ValueGuidedPolicy sums over H (policies.py:262-263)."""

    def __init__(self, w):
        super().__init__()
        self.w = torch.nn.Parameter(torch.as_tensor(w))

    def forward(self, obs):
        return torch.tanh(obs @ self.w)


def residual(c_phys, P):
    return float(torch.mean((c_phys - c_phys @ P) ** 2))


def concat_physical(pol, x):
    s = pol.unnormalize_states(x[:, :, :pol.state_dim])
    a = pol.unnormalize_actions(x[:, :, pol.observation_dim:])
    s = torch.cat([s, s[:, -1:, :]], dim=1)
    return torch.cat([s.reshape(x.shape[0], -1), a.reshape(x.shape[0], -1)], dim=1)


def make_case(name, ref):
    c = helpers.CASES[name]
    T, S, B, H = helpers.case_T(c), c["S"], c["B"], c["H"]
    sd, _ = helpers.make_state_dict(c)
    torch.manual_seed(0)
    net = ref.TemporalUnet(T, dim=c["dim"], dim_mults=c["mults"])
    dif = ref.GaussianDiffusion(net, horizon=H, observation_dim=c["n"], action_dim=c["m"], n_timesteps=S,
                                beta_schedule=c["beta"])
    ref_keys = list(dif.state_dict().keys())
    assert ref_keys == list(sd.keys()), "state_dict key order/layout differs from the reference"
    # the schedule buffers come from the reference's own constructor; everything else from the seeded fill
    own = {k: torch.from_numpy(v) for k, v in sd.items()}
    for k in ref_keys:
        if not k.startswith("model."):
            assert torch.equal(dif.state_dict()[k], own[k]), "schedule buffer %s differs from the reference's" % k
    dif.load_state_dict(own, strict=True)
    dif.eval()
    out = {"case": np.array(helpers.case_json(c)), "keys": np.array("\n".join(ref_keys))}
    x_init, z, start, goal = helpers.noise_inputs(c)
    out.update(x_init=x_init, noise=z, start=start, goal=goal)

    with torch.no_grad():
        # ---- raw U-Net outputs: uniform timesteps and a per-row timestep vector
        xt = torch.from_numpy(x_init)
        steps = sorted({0, S // 2, S - 1})
        out["unet_steps"] = np.array(steps, dtype=np.int64)
        out["unet_eps"] = np.stack([net(xt, torch.full((B,), i, dtype=torch.long)).numpy() for i in steps])
        t_rows = (np.arange(B) * 3 + 1) % S
        out["unet_t_rows"] = t_rows.astype(np.int64)
        out["unet_eps_rows"] = net(xt, torch.from_numpy(t_rows).long()).numpy()

        # ---- GaussianDiffusion.p_sample_loop, traced
        trace = []
        with injected_noise(x_init, z):
            x = torch.randn((B, H, T))
            for i in reversed(range(S)):
                x = dif.p_sample(x, torch.full((B,), i, dtype=torch.long))
                trace.append(x.numpy().copy())
        out["trace_plain"] = np.stack(trace)
        with injected_noise(x_init, z):
            out["final_plain"] = dif.p_sample_loop((B, H, T)).numpy()
        assert np.array_equal(out["final_plain"], out["trace_plain"][-1])

        # ---- GuidedPolicy.sample_loop with start + goal inpainting
        nz = helpers.normalizer(c)
        pol = ref.GuidedPolicy(dif, nz)
        cond = {0: torch.from_numpy(start)[None], H - 1: torch.from_numpy(goal)[None]}
        with injected_noise(x_init, z):
            out["final_cond"] = pol.sample_loop(batch_size=B, conditions=cond).numpy()
        trace = []
        with injected_noise(x_init, z):
            x = pol.apply_conditions(torch.randn((B, H, T)), cond)
            for i in reversed(range(S)):
                x = pol.p_sample_with_guidance(x, torch.full((B,), i, dtype=torch.long), cond)
                trace.append(x.numpy().copy())
        out["trace_cond"] = np.stack(trace)
        assert np.array_equal(out["final_cond"], out["trace_cond"][-1])

    # ---- ValueGuidedPolicy (autograd guidance), start condition only
    vm = ValueModel(helpers.value_weights(c))
    vpol = ref.ValueGuidedPolicy(dif, nz, vm, guide_weight=0.7)
    cond0 = {0: torch.from_numpy(start)[None]}
    trace, grads = [], []
    with injected_noise(x_init, z):
        x = vpol.apply_conditions(torch.randn((B, H, T)), cond0)
        for i in reversed(range(S)):
            x = vpol.p_sample_with_guidance(x, torch.full((B,), i, dtype=torch.long), cond0)
            trace.append(x.detach().numpy().copy())
    out["trace_value"] = np.stack(trace)
    out["value_guide_weight"] = np.array(0.7, dtype=np.float32)

    # ---- dynamics: reference builders
    dyn = helpers.dynamics(c)
    if len(dyn) == 2:
        A, Bm = dyn
    else:
        with contextlib.redirect_stdout(io.StringIO()):
            A, Bm = ref.fit_linear_dynamics(dyn[2], dyn[3], dyn[4])
        out["A_true"], out["B_true"] = dyn[0], dyn[1]
    out["A"], out["Bm"] = np.asarray(A, dtype=np.float64), np.asarray(Bm, dtype=np.float64)
    with contextlib.redirect_stdout(io.StringIO()):
        builder = ref.ProjectionMatrixBuilder(A, Bm, c["n"], c["m"])
        P = builder.get_projection_matrix(H)
    if P.shape[0] <= 256:
        out["P"] = P.numpy()
    out["P_diag"] = torch.diagonal(P).numpy().copy()
    out["P_row0"] = P[0].numpy().copy()
    out["P_fro"] = np.array(float(torch.linalg.norm(P.double())))

    # ---- DynamicsAwarePolicy: apply_projection alone, then the composition of SURVEY.md 8(c)
    with contextlib.redirect_stdout(io.StringIO()):
        dpol = ref.DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=c["n"],
                                       observation_dim=c["n"], action_dim=c["m"], horizon=H,
                                       projection_schedule=c["proj_schedule"], projection_strength=c["strength"])
    with torch.no_grad():
        out["alphas"] = np.array([dpol._get_projection_alpha(i) for i in range(S)], dtype=np.float64)
        out["proj_only"] = np.stack([dpol.apply_projection(torch.from_numpy(x_init), i).numpy() for i in range(S)])
        for order, key in ((False, "dyn"), (True, "dyn_inpaint_first")):
            trace, res = [], []
            with injected_noise(x_init, z):
                x = dpol.apply_conditions(torch.randn((B, H, T)), cond0)
                for i in reversed(range(S)):
                    t = torch.full((B,), i, dtype=torch.long)
                    if order:    # denoise -> inpaint -> project
                        x = dpol.p_sample_with_guidance(x, t, cond0)
                        x = dpol.apply_projection(x, i)
                    else:        # denoise -> project -> inpaint (default)
                        x = dpol.p_sample_with_guidance(x, t, None)
                        x = dpol.apply_projection(x, i)
                        x = dpol.apply_conditions(x, cond0)
                    trace.append(x.numpy().copy())
                    res.append(residual(concat_physical(dpol, x), P))
            out["trace_" + key] = np.stack(trace)
            out["residual_" + key] = np.array(res, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-10s keys=%d  |x_0|=%.4f  residual %.3e -> %.3e" % (
        name, len(ref_keys), float(np.abs(out["trace_dyn"][-1]).mean()), out["residual_dyn"][0], out["residual_dyn"][-1]))


def make_full_case(name, ref):
    """Light golden file of a full-width architecture (helpers.FULL_CASES): raw U-Net outputs (uniform and per-row
    timesteps) and the dynamics-aware composition traced per step, with its residuals.  P itself (up to 2183^2) is
    not stored: tests rebuild it from the stored (A, B) and check it against P_diag / P_row0 / P_fro."""
    c = helpers.FULL_CASES[name]
    T, S, B, H = helpers.case_T(c), c["S"], c["B"], c["H"]
    sd, _ = helpers.make_state_dict(c)
    torch.manual_seed(0)
    net = ref.TemporalUnet(T, dim=c["dim"], dim_mults=c["mults"])
    dif = ref.GaussianDiffusion(net, horizon=H, observation_dim=c["n"], action_dim=c["m"], n_timesteps=S,
                                beta_schedule=c["beta"])
    ref_keys = list(dif.state_dict().keys())
    assert ref_keys == list(sd.keys()), "state_dict key order/layout differs from the reference"
    dif.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    dif.eval()
    out = {"case": np.array(helpers.case_json(c)), "n_keys": np.array(len(ref_keys)),
           "n_params": np.array(sum(p.numel() for p in net.parameters()))}
    x_init, z, start, goal = helpers.noise_inputs(c)
    out.update(x_init=x_init, noise=z, start=start, goal=goal)
    nz = helpers.normalizer(c)
    cond0 = {0: torch.from_numpy(start)[None]}
    with torch.no_grad():
        xt = torch.from_numpy(x_init)
        steps = sorted({0, S - 1})
        out["unet_steps"] = np.array(steps, dtype=np.int64)
        out["unet_eps"] = np.stack([net(xt, torch.full((B,), i, dtype=torch.long)).numpy() for i in steps])
        t_rows = (np.arange(B) * 2 + 1) % S
        out["unet_t_rows"] = t_rows.astype(np.int64)
        out["unet_eps_rows"] = net(xt, torch.from_numpy(t_rows).long()).numpy()
        dyn = helpers.dynamics(c)
        with contextlib.redirect_stdout(io.StringIO()):
            A, Bm = ref.fit_linear_dynamics(dyn[2], dyn[3], dyn[4])
            P = ref.ProjectionMatrixBuilder(A, Bm, c["n"], c["m"]).get_projection_matrix(H)
            dpol = ref.DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=c["n"],
                                           observation_dim=c["n"], action_dim=c["m"], horizon=H,
                                           projection_schedule=c["proj_schedule"], projection_strength=c["strength"])
        out["A"], out["Bm"] = np.asarray(A, dtype=np.float64), np.asarray(Bm, dtype=np.float64)
        out["P_diag"] = torch.diagonal(P).numpy().copy()
        out["P_row0"] = P[0].numpy().copy()
        out["P_fro"] = np.array(float(torch.linalg.norm(P.double())))
        out["alphas"] = np.array([dpol._get_projection_alpha(i) for i in range(S)], dtype=np.float64)
        trace, res = [], []
        with injected_noise(x_init, z):
            x = dpol.apply_conditions(torch.randn((B, H, T)), cond0)
            for i in reversed(range(S)):
                x = dpol.p_sample_with_guidance(x, torch.full((B,), i, dtype=torch.long), None)
                x = dpol.apply_projection(x, i)
                x = dpol.apply_conditions(x, cond0)
                trace.append(x.numpy().copy())
                res.append(residual(concat_physical(dpol, x), P))
        out["trace_dyn"] = np.stack(trace)
        out["residual_dyn"] = np.array(res, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-12s params=%d  |eps|=%.4f  residual %.3e -> %.3e" % (
        name, int(out["n_params"]), float(np.abs(out["unet_eps"]).mean()), out["residual_dyn"][0], out["residual_dyn"][-1]))


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = ref_shim.load()
    for name in (sys.argv[1:] or list(helpers.CASES) + list(helpers.FULL_CASES)):
        (make_full_case if name in helpers.FULL_CASES else make_case)(name, ref)
