"""The oracle (numpy restatement) against every golden vector produced by the live reference
(tests/golden/make_golden.py).  CPU only.  fp64 oracle vs the reference's fp32 torch: 2e-5 relative L2
per step when teacher-forced; traces are compared teacher-forced because the clamp at +-1
(diffusion.py:199-200) makes free-running drift non-smooth."""
import numpy as np
import pytest

import helpers
from oracle.diffusion import make_buffers, BUFFER_NAMES, apply_conditions

TOL = 2e-5
CASE_NAMES = list(helpers.CASES)


@pytest.fixture(scope="module", params=CASE_NAMES)
def setup(request):
    c = helpers.CASES[request.param]
    sd, _ = helpers.make_state_dict(c)
    g = helpers.load_golden(request.param)
    unet, dif, proj, P = helpers.build_oracle(c, sd)
    return c, sd, g, unet, dif, proj, P


def test_state_dict_layout(setup):
    c, sd, g, *_ = setup
    assert str(g["keys"]).split("\n") == list(sd.keys())
    assert str(g["case"]) == helpers.case_json(c)


def test_schedule_buffers_bitwise(setup):
    """diffusion.py:96-128 restated in numpy fp32 reproduces the reference's registered buffers."""
    c, sd, *_ = setup
    b = make_buffers(c["S"], c["beta"], dtype=np.float32)
    for k in BUFFER_NAMES:
        np.testing.assert_allclose(b[k], sd[k], rtol=3e-6, atol=1e-7, err_msg=k)


def test_unet_forward(setup):
    c, sd, g, unet, *_ = setup
    for i, want in zip(g["unet_steps"], g["unet_eps"]):
        got = unet.forward(g["x_init"], np.full((c["B"],), i))
        assert helpers.rel_l2(got, want) < TOL
    got = unet.forward(g["x_init"], g["unet_t_rows"])
    assert helpers.rel_l2(got, g["unet_eps_rows"]) < TOL


def _teacher_forced(trace, x_init, step_fn, S):
    worst = 0.0
    for k, i in enumerate(reversed(range(S))):
        x_in = x_init if k == 0 else trace[k - 1]
        got = step_fn(np.array(x_in, dtype=np.float64), i, k)
        worst = max(worst, helpers.rel_l2(got, trace[k]))
    return worst


def test_p_sample_loop_trace(setup):
    c, sd, g, unet, dif, *_ = setup
    worst = _teacher_forced(g["trace_plain"], g["x_init"], lambda x, i, k: dif.p_sample(x, i, g["noise"][k]), c["S"])
    assert worst < TOL


def test_conditioned_trace(setup):
    c, sd, g, unet, dif, *_ = setup
    cond = {0: g["start"], c["H"] - 1: g["goal"]}
    x0 = apply_conditions(np.array(g["x_init"], dtype=np.float64), cond)
    worst = _teacher_forced(g["trace_cond"], x0, lambda x, i, k: dif.guided_step(x, i, g["noise"][k], cond), c["S"])
    assert worst < TOL


def test_value_guided_trace(setup):
    c, sd, g, unet, dif, *_ = setup
    w = helpers.value_weights(c).astype(np.float64)
    n = c["n"]

    def grad(x):
        # d/dx sum_{b,h} tanh(obs . w) = (1 - tanh^2) w on the observation dims, 0 on the action dims
        s = np.tanh(x[:, :, :n] @ w)
        gr = np.zeros_like(x)
        gr[:, :, :n] = (1 - s ** 2)[:, :, None] * w
        return gr

    cond = {0: g["start"]}
    x0 = apply_conditions(np.array(g["x_init"], dtype=np.float64), cond)
    gw = float(g["value_guide_weight"])
    worst = _teacher_forced(g["trace_value"], x0,
                            lambda x, i, k: dif.guided_step(x, i, g["noise"][k], cond, grad(x), gw), c["S"])
    assert worst < TOL


def test_projection_matrix(setup):
    c, sd, g, unet, dif, proj, P = setup
    np.testing.assert_allclose(np.diagonal(P), g["P_diag"], atol=2e-5)
    np.testing.assert_allclose(P[0], g["P_row0"], atol=2e-5)
    assert abs(np.linalg.norm(P.astype(np.float64)) - float(g["P_fro"])) < 1e-3
    if "P" in g:
        np.testing.assert_allclose(P, g["P"], atol=2e-5)
    # the reference's own known-answer: idempotence (projection.py:110-117,132-133)
    assert np.allclose(P @ P, P, atol=1e-4)
    if "A_true" in g:   # fit_linear_dynamics (data_driven.py:75-134) recovers the generating system
        dyn = helpers.dynamics(c)
        from oracle.projection import fit_linear_dynamics
        A, Bm = fit_linear_dynamics(dyn[2], dyn[3], dyn[4])
        np.testing.assert_allclose(A, g["A"], atol=1e-9)
        np.testing.assert_allclose(Bm, g["Bm"], atol=1e-9)
        assert np.abs(A - g["A_true"]).max() < 1e-3


def test_projection_alpha_and_apply(setup):
    c, sd, g, unet, dif, proj, P = setup
    for i in range(c["S"]):
        assert abs(proj.alpha(i) - g["alphas"][i]) < 1e-7
        got = proj.apply(np.array(g["x_init"], dtype=np.float64), i)
        assert helpers.rel_l2(got, g["proj_only"][i]) < TOL


@pytest.mark.parametrize("order", ["dyn", "dyn_inpaint_first"])
def test_dynamics_aware_trace(setup, order):
    c, sd, g, unet, dif, proj, P = setup
    cond = {0: g["start"]}
    x0 = apply_conditions(np.array(g["x_init"], dtype=np.float64), cond)

    def step(x, i, k):
        if order == "dyn":
            x = dif.guided_step(x, i, g["noise"][k], None)
            return apply_conditions(proj.apply(x, i), cond)
        return proj.apply(dif.guided_step(x, i, g["noise"][k], cond), i)

    trace = g["trace_" + order]
    worst = _teacher_forced(trace, x0, step, c["S"])
    assert worst < TOL
    # dynamics residual (losses/__init__.py:161-186) of the reference's own x_i, recomputed by the oracle
    for k in range(c["S"]):
        want = g["residual_" + order][k]
        assert abs(proj.residual(trace[k]) - want) <= 1e-4 * max(want, 1e-3)


def test_free_running_loop_matches_when_unclamped_drift_is_small(setup):
    """Whole sample_loop, free-running, against the reference's final x_0 (looser: drift accumulates)."""
    c, sd, g, unet, dif, proj, P = setup
    x = dif.sample_loop(g["x_init"], g["noise"], conditions={0: g["start"]}, projector=proj)
    assert helpers.rel_l2(x, g["trace_dyn"][-1]) < 1e-3


def test_torch_port_matches_reference(setup):
    """oracle/torch_port.py (the timed CPU baseline and the stock-bf16 calibration) reproduces the reference's
    own free-running dynamics-aware trace: same ATen ops in the same order, so fp32 round-off level."""
    import torch
    from oracle import torch_port
    c, sd, g, unet, dif, proj, P = setup
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    nz = helpers.normalizer(c)
    nzt = tuple(torch.from_numpy(np.asarray(a, dtype=np.float32)) for a in (nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std))
    projector = dict(P=torch.from_numpy(P), alphas=[float(a) for a in g["alphas"]], nz=nzt, n=c["n"], m=c["m"], H=c["H"])
    x = torch_port.sample_loop(tsd, torch.from_numpy(np.array(g["x_init"])), torch.from_numpy(g["noise"]),
                               conditions={0: torch.from_numpy(g["start"])}, projector=projector)
    assert helpers.rel_l2(x.numpy(), g["trace_dyn"][-1]) < 5e-5
    x = torch_port.sample_loop(tsd, torch.from_numpy(np.array(g["x_init"])), torch.from_numpy(g["noise"]),
                               conditions={0: torch.from_numpy(g["start"]), c["H"] - 1: torch.from_numpy(g["goal"])})
    assert helpers.rel_l2(x.numpy(), g["trace_cond"][-1]) < 5e-5
