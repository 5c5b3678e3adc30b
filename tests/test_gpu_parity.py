"""GPU parity: the CUDA path (through the reference-shaped Python classes -> ctypes -> C ABI) against the
committed golden vectors of the live reference and against the numpy oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): per-step relative L2 <= 1e-5 in fp32 mode, <= 1e-2 in bf16 mode,
asserted TEACHER-FORCED (every step starts from the reference's x_i); free-running drift is bounded looser.
"""
import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu

# "bf16" runs the throughput kernels (conv_t3 / conv_tc) at the test batch sizes, "bf16-latency" the small-batch
# kernels (conv_small, the get_action path): same arithmetic contract, same tolerances.
PRECISIONS = ["fp32", "bf16", "bf16-latency"]
TOL = {"fp32": 1e-5, "bf16": 1e-2, "bf16-latency": 1e-2}
FREE_TOL = {"fp32": 2e-4, "bf16": 5e-2, "bf16-latency": 5e-2}
# The ONE ill-conditioned step: with the cosine schedule beta_{S-1} is clipped to 0.9999 (diffusion.py:41), so
# the first reverse step computes x0 = 100 x - 99.99 eps and the posterior mean has d(mean)/d(eps) = 99.98
# (every other step: < 1.5).  Any bf16 evaluation of eps (ours: 8e-3 relative; stock torch.autocast: 1.1e-2,
# tools/gpu_probe.py) is amplified ~40x there (observed 1.3-2e-2 on x), so bf16 models take the eps of that step from
# the fp32 kernels (GaussianDiffusion.fp32_ill_conditioned_steps, default True; dad_set_fp32_steps) and EVERY step is
# held to BASELINE.json's 1e-2.  The eps error of the bf16 kernels themselves must not exceed stock autocast-bf16
# PyTorch's on the same input (test_bf16_unet_no_worse_than_stock_autocast).
ILL_TOL = 1e-2
# Raw U-Net output (eps) in bf16: the tolerance of BASELINE.json is on x per step; eps itself carries the bf16
# rounding of ~35 layers (ours 0.8-1.05e-2, stock torch.autocast 1.1-1.3e-2 on the same inputs) and is held to
# 1.5e-2 here and to <= 1.1x stock autocast in test_bf16_unet_no_worse_than_stock_autocast.
EPS_TOL = {"fp32": 1e-5, "bf16": 1.5e-2, "bf16-latency": 1.5e-2}
CASE_NAMES = list(helpers.CASES)


def step_tols(c, sd, precision):
    """Per-step tolerance, index k = loop iteration (step i = S-1-k)."""
    tol = []
    for i in reversed(range(c["S"])):
        amp = float(sd["posterior_mean_coef1"][i] * sd["sqrt_recipm1_alphas_cumprod"][i])
        tol.append(ILL_TOL if (precision.startswith("bf16") and amp > 10.0) else TOL[precision])
    return tol


def worst_excess(errs, tols):
    return max(e / t for e, t in zip(errs, tols))


def _dev():
    return torch.device("cuda", 0)


_cache = {}


def models(name, precision):
    """(case, golden, our GaussianDiffusion on cuda:0 with the case's weights)."""
    key = (name, precision)
    if key not in _cache:
        from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion
        c = helpers.CASES[name]
        sd, _ = helpers.make_state_dict(c)
        prec, _, mode = precision.partition("-")
        net = TemporalUnet(helpers.case_T(c), dim=c["dim"], dim_mults=c["mults"], precision=prec, max_batch=64,
                           latency_max_batch=(8 if mode == "latency" else 0) if prec == "bf16" else None)
        dif = GaussianDiffusion(net, horizon=c["H"], observation_dim=c["n"], action_dim=c["m"], n_timesteps=c["S"],
                                beta_schedule=c["beta"])
        dif.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        dif.to(_dev())
        _cache.clear()      # one engine set alive at a time keeps device memory small
        _cache[key] = (c, helpers.load_golden(name), dif, sd)
    return _cache[key]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(_dev())


def dyn_policy(c, dif, P):
    from dynamics_aware_diffusion_b200 import DynamicsAwarePolicy
    return DynamicsAwarePolicy(dif, projection_matrix=torch.from_numpy(P), normalizer=helpers.normalizer(c),
                               state_dim=c["n"], observation_dim=c["n"], action_dim=c["m"], horizon=c["H"],
                               projection_schedule=c["proj_schedule"], projection_strength=c["strength"])


def case_P(c, g):
    if "P" in g:
        return g["P"]
    from dynamics_aware_diffusion_b200 import ProjectionMatrixBuilder
    return ProjectionMatrixBuilder(g["A"], g["Bm"], c["n"], c["m"]).get_projection_matrix(c["H"]).numpy()


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_unet_forward(name, precision):
    c, g, dif, _ = models(name, precision)
    x = cu(g["x_init"])
    for i, want in zip(g["unet_steps"], g["unet_eps"]):
        got = dif.model(x, torch.full((c["B"],), int(i), device=x.device, dtype=torch.long))
        assert helpers.rel_l2(got.cpu().numpy(), want) < EPS_TOL[precision], "step %d" % i
    got = dif.model(x, cu(g["unet_t_rows"]))
    assert helpers.rel_l2(got.cpu().numpy(), g["unet_eps_rows"]) < EPS_TOL[precision]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_p_sample_teacher_forced(name, precision, monkeypatch):
    """GaussianDiffusion.p_sample per step, noise injected through torch.randn_like like the golden generator."""
    c, g, dif, sd = models(name, precision)
    S = c["S"]
    errs = []
    for k, i in enumerate(reversed(range(S))):
        x_in = g["x_init"] if k == 0 else g["trace_plain"][k - 1]
        z = cu(g["noise"][k])
        monkeypatch.setattr(torch, "randn_like", lambda t, **kw: z)
        got = dif.p_sample(cu(x_in), torch.full((c["B"],), i, device=_dev(), dtype=torch.long))
        errs.append(helpers.rel_l2(got.cpu().numpy(), g["trace_plain"][k]))
    assert worst_excess(errs, step_tols(c, sd, precision)) < 1.0, errs


@pytest.mark.parametrize("name", CASE_NAMES)
def test_bf16_unet_no_worse_than_stock_autocast(name):
    """Calibration of the bf16 mode: the hand-written tcgen05 path deviates from the fp32 reference no more than
    stock torch.autocast(bf16) running the same network (oracle/torch_port.py) on this GPU."""
    from oracle import torch_port
    c, g, dif, sd = models(name, "bf16")
    x = cu(g["x_init"])
    w = {k[len("model."):]: cu(v) for k, v in sd.items() if k.startswith("model.")}
    for i, want in zip(g["unet_steps"], g["unet_eps"]):
        t = torch.full((c["B"],), int(i), device=_dev(), dtype=torch.long)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            e_stock = helpers.rel_l2(torch_port.unet_forward(w, x, t).float().cpu().numpy(), want)
        e_ours = helpers.rel_l2(dif.model(x, t).cpu().numpy(), want)
        assert e_ours <= 1.1 * e_stock, (e_ours, e_stock)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_p_mean_variance(name, precision):
    c, g, dif, sd = models(name, precision)
    _, odif, _, _ = helpers.build_oracle(c, sd)
    i = c["S"] // 2
    t = torch.full((c["B"],), i, device=_dev(), dtype=torch.long)
    mean, logvar = dif.p_mean_variance(cu(g["x_init"]), t)
    want_mean, want_lv, _ = odif.p_mean_variance(np.array(g["x_init"], dtype=np.float64), i)
    assert helpers.rel_l2(mean.cpu().numpy(), want_mean) < TOL[precision]
    assert tuple(logvar.shape) == (c["B"], 1, 1)
    assert abs(float(logvar[0, 0, 0]) - float(want_lv)) < 1e-5 * max(1.0, abs(float(want_lv)))


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_conditioned_and_value_guided_steps(name, precision, monkeypatch):
    from dynamics_aware_diffusion_b200 import GuidedPolicy, ValueGuidedPolicy
    c, g, dif, sd = models(name, precision)
    S, H = c["S"], c["H"]
    tols = step_tols(c, sd, precision)
    nz = helpers.normalizer(c)
    pol = GuidedPolicy(dif, nz)
    cond = {0: cu(g["start"])[None], H - 1: cu(g["goal"])[None]}
    x0 = np.array(g["x_init"])
    x0[:, 0], x0[:, H - 1] = g["start"], g["goal"]
    errs = []
    for k, i in enumerate(reversed(range(S))):
        x_in = x0 if k == 0 else g["trace_cond"][k - 1]
        z = cu(g["noise"][k])
        monkeypatch.setattr(torch, "randn_like", lambda t, **kw: z)
        got = pol.p_sample_with_guidance(cu(x_in), torch.full((c["B"],), i, device=_dev(), dtype=torch.long), cond)
        errs.append(helpers.rel_l2(got.cpu().numpy(), g["trace_cond"][k]))
        # inpainting is exact
        assert np.array_equal(got[:, 0].cpu().numpy(), np.broadcast_to(g["start"], (c["B"], helpers.case_T(c))))
    assert worst_excess(errs, tols) < 1.0, errs

    class ValueModel(torch.nn.Module):
        def __init__(self, w):
            super().__init__()
            self.w = torch.nn.Parameter(w)

        def forward(self, obs):
            return torch.tanh(obs @ self.w)

    vpol = ValueGuidedPolicy(dif, nz, ValueModel(cu(helpers.value_weights(c))), guide_weight=float(g["value_guide_weight"]))
    cond0 = {0: cu(g["start"])[None]}
    x0 = np.array(g["x_init"])
    x0[:, 0] = g["start"]
    errs = []
    for k, i in enumerate(reversed(range(S))):
        x_in = x0 if k == 0 else g["trace_value"][k - 1]
        z = cu(g["noise"][k])
        monkeypatch.setattr(torch, "randn_like", lambda t, **kw: z)
        got = vpol.p_sample_with_guidance(cu(x_in), torch.full((c["B"],), i, device=_dev(), dtype=torch.long), cond0)
        errs.append(helpers.rel_l2(got.cpu().numpy(), g["trace_value"][k]))
    assert worst_excess(errs, tols) < 1.0, errs


@pytest.mark.parametrize("precision", ["fp32", "bf16-latency"])
@pytest.mark.parametrize("name", ["tiny", "pointmaze", "cheetah_s"])
def test_value_guided_loop_graph_equals_host_loop(name, precision, monkeypatch):
    """ValueGuidedPolicy.sample_loop: the captured step (our kernels + the value model's autograd in one CUDA
    graph, step index on the device) against the per-step host loop and against the reference's free-running
    value-guided trace; a guide_fn that cannot be captured falls back to the host loop."""
    from dynamics_aware_diffusion_b200 import ValueGuidedPolicy
    c, g, dif, sd = models(name, precision)

    class ValueModel(torch.nn.Module):
        def __init__(self, w):
            super().__init__()
            self.w = torch.nn.Parameter(w)

        def forward(self, obs):
            return torch.tanh(obs @ self.w)

    vm = ValueModel(cu(helpers.value_weights(c)))
    cond0 = {0: cu(g["start"])[None]}
    x_init = cu(g["x_init"])
    monkeypatch.setattr(torch, "randn", lambda *a, **k: x_init.clone())
    outs = {}
    for mode in ("graph", "host"):
        vpol = ValueGuidedPolicy(dif, helpers.normalizer(c), vm, guide_weight=float(g["value_guide_weight"]))
        vpol.capture_guidance = mode == "graph"
        x, trace = vpol.sample_loop(batch_size=c["B"], conditions=cond0, noise=cu(g["noise"]), return_trace=True)
        if mode == "graph":
            assert vpol.capture_guidance, "capture failed: %s" % vpol._capture_error
            # the cached captured step serves the next call (other conditions, Philox noise)
            x2 = vpol.sample_loop(batch_size=c["B"], conditions={0: cu(g["goal"])[None]}, seed=3)
            assert bool(torch.isfinite(x2).all()) and bool((x2[:, 0] == cu(g["goal"])).all())
            x3 = vpol.sample_loop(batch_size=c["B"], conditions=cond0, noise=cu(g["noise"]))
            assert torch.equal(x3, x)
        outs[mode] = (x, trace)
        assert torch.equal(trace[-1], x)
    assert helpers.rel_l2(outs["graph"][1].cpu().numpy(), outs["host"][1].cpu().numpy()) < 1e-6
    assert helpers.rel_l2(outs["graph"][1].cpu().numpy(), g["trace_value"]) < FREE_TOL[precision]

    # not capturable: the guide synchronises with the host
    vpol = ValueGuidedPolicy(dif, helpers.normalizer(c), vm, guide_weight=float(g["value_guide_weight"]))
    inner = vpol.guide_fn
    vpol.guide_fn = lambda x, t: inner(x, t) * float(t[0].item() >= 0)
    x = vpol.sample_loop(batch_size=c["B"], conditions=cond0, noise=cu(g["noise"]))
    assert not vpol.capture_guidance and vpol._capture_error
    assert helpers.rel_l2(x.cpu().numpy(), outs["host"][0].cpu().numpy()) < 1e-6


@pytest.mark.parametrize("name", CASE_NAMES)
def test_apply_projection(name):
    """DynamicsAwarePolicy.apply_projection (one affine map on the device) vs the reference's 15-op chain."""
    c, g, dif, _ = models(name, "fp32")
    pol = dyn_policy(c, dif, case_P(c, g))
    for i in range(c["S"]):
        assert abs(pol._get_projection_alpha(i) - g["alphas"][i]) < 1e-7
        got = pol.apply_projection(cu(g["x_init"]), i)
        assert helpers.rel_l2(got.cpu().numpy(), g["proj_only"][i]) < 1e-5


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("order", ["dyn", "dyn_inpaint_first"])
@pytest.mark.parametrize("name", CASE_NAMES)
def test_dynamics_aware_loop(name, order, precision):
    """The headline path: DynamicsAwarePolicy.sample_loop = ONE dad_sample call (graph replays of U-Net + fused
    step/projection/inpaint kernel).  Teacher-forced per step via 1-step loops; then the free-running loop with
    its per-step trace; then the dynamics residual of the result."""
    from dynamics_aware_diffusion_b200 import _native as N
    c, g, dif, sd = models(name, precision)
    P = case_P(c, g)
    pol = dyn_policy(c, dif, P)
    pol.project_after_inpaint = order == "dyn_inpaint_first"
    S, B = c["S"], c["B"]
    trace = g["trace_" + order]
    cond0 = {0: cu(g["start"])[None]}
    eng = pol._engine(_dev())
    flags = pol._loop_flags(eng) | N.FLAG_CONDITIONS
    eng.set_conditions(cond0, B)
    x0 = np.array(g["x_init"])
    x0[:, 0] = g["start"]
    errs = []
    for k, i in enumerate(reversed(range(S))):
        x = cu(x0 if k == 0 else trace[k - 1])
        eps = dif.eps_engine(eng, i).unet_forward(x, step=i)
        eng.step(x, eps, i, noise=cu(g["noise"][k]), flags=flags)
        errs.append(helpers.rel_l2(x.cpu().numpy(), trace[k]))
    assert worst_excess(errs, step_tols(c, sd, precision)) < 1.0, ("teacher-forced", errs)
    # free-running, whole loop on the device with injected noise
    torch.manual_seed(0)
    real = torch.randn
    try:
        torch.randn = lambda *a, **k: cu(g["x_init"])
        out, tr = pol.sample_loop(batch_size=B, conditions=cond0, noise=cu(g["noise"]), return_trace=True)
    finally:
        torch.randn = real
    assert tuple(tr.shape) == (S, B, c["H"], helpers.case_T(c))
    assert torch.equal(tr[-1], out)
    assert helpers.rel_l2(out.cpu().numpy(), trace[-1]) < FREE_TOL[precision], "free-running"
    # dynamics residual ||tau - tau P||^2 (losses/__init__.py:161-186) of OUR trajectories vs the reference's
    _, _, oproj, _ = helpers.build_oracle(c, sd)
    for k in (0, S - 1):
        want = g["residual_" + order][k]
        got = oproj.residual(tr[k].cpu().numpy())
        assert abs(got - want) <= TOL[precision] * 5 * max(want, 1e-3) + (0 if precision == "fp32" else 1e-3 * want)


@pytest.mark.parametrize("name", ["tiny", "cheetah_s"])
def test_sample_host_c_abi(name):
    """dad_sample_host: host buffers in/out through the C ABI (the `e2e` bench leg)."""
    from dynamics_aware_diffusion_b200 import _native as N
    c, g, dif, _ = models(name, "fp32")
    pol = dyn_policy(c, dif, case_P(c, g))
    eng = pol._engine(_dev())
    flags = pol._loop_flags(eng) | N.FLAG_CONDITIONS
    eng.set_conditions({0: cu(g["start"])[None]}, c["B"])
    x = torch.from_numpy(np.array(g["x_init"])).pin_memory()
    x[:, 0] = torch.from_numpy(g["start"])
    z = torch.from_numpy(np.array(g["noise"])).pin_memory()
    eng.sample_host(x, c["S"], noise_seq_host=z, flags=flags)
    assert helpers.rel_l2(x.numpy(), g["trace_dyn"][-1]) < FREE_TOL["fp32"]


def test_raw_c_abi_sampling():
    """The C ABI with nothing but ctypes (include/dad_b200.h as a non-PyTorch host would bind it):
    dad_create -> dad_load_weights -> dad_set_schedule -> dad_set_conditions -> dad_sample_host, checked against
    the reference's conditioned trace; then the error convention."""
    import ctypes
    from dynamics_aware_diffusion_b200 import _native as N
    c = helpers.CASES["tiny"]
    g = helpers.load_golden("tiny")
    sd, _ = helpers.make_state_dict(c)
    L = N.lib()
    cfg = N.DadConfig(abi_version=N.DAD_ABI_VERSION, device=0, precision=N.PRECISION_FP32, transition_dim=6, dim=c["dim"],
                      n_levels=len(c["mults"]), kernel_size=5, time_dim=0, horizon=c["H"], n_timesteps=c["S"],
                      predict_epsilon=1, clip_denoised=1, max_batch=8)
    for i, m in enumerate(c["mults"]):
        cfg.dim_mults[i] = m
    h = ctypes.c_void_p()
    assert L.dad_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    try:
        unet = {k[len("model."):]: np.ascontiguousarray(v, dtype=np.float32) for k, v in sd.items() if k.startswith("model.")}
        names = [k.encode() for k in unet]
        arr = (N.DadTensor * len(unet))()
        for i, (nm, v) in enumerate(zip(names, unet.values())):
            arr[i].name, arr[i].data, arr[i].numel = nm, v.ctypes.data, v.size        # HOST pointers are accepted
        assert L.dad_load_weights(h, arr, len(unet)) == 0
        bufs = [np.ascontiguousarray(sd[k], dtype=np.float32) for k in
                ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                 "posterior_mean_coef2", "posterior_log_variance_clipped")]
        assert L.dad_set_schedule(h, *[ctypes.c_void_p(b.ctypes.data) for b in bufs], c["S"]) == 0
        hidx = (ctypes.c_int32 * 2)(0, -1)                                        # python-style negative index = goal
        vals = np.ascontiguousarray(np.stack([g["start"], g["goal"]]), dtype=np.float32)
        assert L.dad_set_conditions(h, hidx, ctypes.c_void_p(vals.ctypes.data), 2, 0, c["B"]) == 0
        x = np.array(g["x_init"], dtype=np.float32)
        z = np.ascontiguousarray(g["noise"], dtype=np.float32)
        rc = L.dad_sample_host(h, ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(z.ctypes.data), 0, 0, c["B"], c["S"],
                               N.FLAG_CONDITIONS)
        assert rc == 0, L.dad_last_error(h)
        assert helpers.rel_l2(x, g["trace_cond"][-1]) < FREE_TOL["fp32"]
        # errors: negative status + message, nothing thrown
        assert L.dad_sample_host(h, ctypes.c_void_p(x.ctypes.data), None, 0, 0, c["B"], c["S"] + 1, 0) == N.ERR_INVALID
        assert b"n_steps" in L.dad_last_error(h)
        assert L.dad_step(h, None, None, None, None, 0.0, 0, 0, 0, 0, 1, None) == N.ERR_INVALID
        info = N.DadInfo()
        assert L.dad_get_info(h, ctypes.byref(info)) == 0 and info.n_conv_layers > 0 and info.sm_count > 0
        assert L.dad_launch_count(h) > 0
    finally:
        assert L.dad_destroy(h) == 0


@pytest.mark.parametrize("name", ["tiny", "pointmaze", "door_s"])
def test_fp32_ill_conditioned_steps_option(name):
    """`diffusion.fp32_ill_conditioned_steps` (default True): the one reverse step with d(mean)/d(eps) = 99.98 (cosine schedule,
    diffusion.py:41) takes its eps from the fp32 kernels, and then EVERY step of the bf16 mode is inside the 1e-2
    tolerance of BASELINE.json -- the first step is compared teacher-forced (it starts from x_S), the rest through the
    free-running trace.  Noise and trace slots are unchanged by the split."""
    c, g, dif, sd = models(name, "bf16")
    assert dif.ill_conditioned_prefix() == 1
    P = case_P(c, g)
    pol = dyn_policy(c, dif, P)
    cond0 = {0: cu(g["start"])[None]}
    real = torch.randn
    outs = {}
    try:
        torch.randn = lambda *a, **k: cu(g["x_init"])
        for on in (False, True):
            dif.fp32_ill_conditioned_steps = on
            outs[on] = pol.sample_loop(batch_size=c["B"], conditions=cond0, noise=cu(g["noise"]), return_trace=True)
    finally:
        torch.randn = real
        dif.fp32_ill_conditioned_steps = True
    trace = g["trace_dyn"]
    first_off = helpers.rel_l2(outs[False][1][0].cpu().numpy(), trace[0])
    first_on = helpers.rel_l2(outs[True][1][0].cpu().numpy(), trace[0])
    assert first_on < TOL["bf16"], (first_on, first_off)
    # the TF32 sibling: at most half the bf16 error, or already far inside the tolerance
    assert first_on < max(0.5 * first_off, 0.3 * TOL["bf16"]), (first_on, first_off)
    assert torch.equal(outs[True][1][-1], outs[True][0])
    assert helpers.rel_l2(outs[True][0].cpu().numpy(), trace[-1]) < FREE_TOL["bf16"]
    # Philox path: same split, results finite, conditions exact
    x = pol.sample_loop(batch_size=c["B"], conditions=cond0, seed=3)
    assert bool(torch.isfinite(x).all()) and bool((x[:, 0] == cu(g["start"])).all())


def test_fp32_ill_steps_is_a_no_op_for_the_linear_schedule():
    c, g, dif, sd = models("cheetah_s", "bf16")
    assert dif.ill_conditioned_prefix() == 0
    pol = dyn_policy(c, dif, case_P(c, g))
    outs = []
    for on in (False, True):
        dif.fp32_ill_conditioned_steps = on
        torch.manual_seed(0)
        outs.append(pol.sample_loop(batch_size=c["B"], seed=9))
    dif.fp32_ill_conditioned_steps = True
    assert torch.equal(outs[0], outs[1])
