"""CPU tests of the host-side logic: the affine fold of apply_projection, projection schedules, the projector
builder, state_dict round-trips, checkpoint layout, the policies' environment glue, the C ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import helpers
from dynamics_aware_diffusion_b200 import (TemporalUnet, GaussianDiffusion, GuidedPolicy, MPCPolicy,
                                           DynamicsAwarePolicy, ProjectionMatrixBuilder, fit_linear_dynamics,
                                           fold_projection, projection_alphas, _native)

CASE_NAMES = list(helpers.CASES)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_fold_projection_equals_reference_chain(name):
    """y = x + alpha (N x + q) reproduces the reference's apply_projection outputs (golden `proj_only`)."""
    c = helpers.CASES[name]
    g = helpers.load_golden(name)
    if "P" in g:
        P = g["P"]
    else:
        P = ProjectionMatrixBuilder(g["A"], g["Bm"], c["n"], c["m"]).get_projection_matrix(c["H"]).numpy()
    nz = helpers.normalizer(c)
    Nm, q = fold_projection(P, nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std, c["n"], c["m"], c["H"])
    sd, _ = helpers.make_state_dict(c)
    al = projection_alphas(c["S"], c["S"], c["proj_schedule"], c["strength"], sd["betas"])
    np.testing.assert_allclose(al, np.where(g["alphas"] > 0, g["alphas"], 0), atol=1e-7)
    x = g["x_init"].reshape(c["B"], -1).astype(np.float64)
    for i in range(c["S"]):
        y = x + al[i] * (x @ Nm.T + q)
        assert helpers.rel_l2(y.reshape(g["x_init"].shape), g["proj_only"][i]) < 2e-5


@pytest.mark.parametrize("name", CASE_NAMES)
def test_projection_builder_matches_reference(name):
    c = helpers.CASES[name]
    g = helpers.load_golden(name)
    b = ProjectionMatrixBuilder(g["A"], g["Bm"], c["n"], c["m"])
    P = b.get_projection_matrix(c["H"])
    assert P.dtype == torch.float32
    D = (c["H"] + 1) * c["n"] + c["H"] * c["m"]
    assert tuple(P.shape) == (D, D)
    np.testing.assert_allclose(np.diagonal(P.numpy()), g["P_diag"], atol=2e-5)
    np.testing.assert_allclose(P[0].numpy(), g["P_row0"], atol=2e-5)
    assert b.verify_projection(P)
    if "A_true" in g:
        dyn = helpers.dynamics(c)
        A, Bm = fit_linear_dynamics(dyn[2], dyn[3], dyn[4])
        np.testing.assert_allclose(A, g["A"], atol=1e-9)
        np.testing.assert_allclose(Bm, g["Bm"], atol=1e-9)


def test_projection_schedules():
    betas = np.linspace(1e-4, 0.02, 10).astype(np.float32)
    assert np.allclose(projection_alphas(10, 10, "constant", 0.5), 0.5)
    assert np.allclose(projection_alphas(10, 10, "linear", 1.0), 1 - np.arange(10) / 10)
    assert np.allclose(projection_alphas(10, 10, "quadratic", 2.0), 2 * (1 - np.arange(10) / 10) ** 2)
    assert np.allclose(projection_alphas(10, 10, "noise_schedule", 1.0, betas), np.sqrt(1 - betas), atol=1e-7)
    # truncated loops (evaluate.py:351-353): progress uses the policy's n_timesteps; alpha <= 0 -> no projection
    a = projection_alphas(10, 5, "linear", 1.0)
    assert a[5] == 0 and np.all(a[6:] == 0) and a[0] == 1
    with pytest.raises(ValueError):
        projection_alphas(10, 10, "cubic", 1.0)


def test_state_dict_roundtrip_and_checkpoint_layout(tmp_path):
    """Keys/shapes are the reference's (golden `keys`); a training checkpoint dict (utils/training.py:193-211)
    loads through model_state_dict."""
    c = helpers.CASES["pointmaze"]
    sd, dif = helpers.make_state_dict(c)
    g = helpers.load_golden("pointmaze")
    assert list(sd.keys()) == str(g["keys"]).split("\n")
    assert sum(v.size for k, v in sd.items() if k.startswith("model.")) == 15_860_486
    ckpt = {"epoch": 3, "global_step": 10, "model_state_dict": dif.state_dict(),
            "config": {"horizon": 32, "observation_dim": 4, "action_dim": 2, "n_timesteps": c["S"], "beta_schedule": "cosine"}}
    path = tmp_path / "checkpoint_step_10.pt"
    torch.save(ckpt, path)
    net = TemporalUnet(6, dim=128, dim_mults=(1, 2, 4))
    dif2 = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=c["S"])
    dif2.load_state_dict(torch.load(path, weights_only=False)["model_state_dict"], strict=True)
    for k, v in dif2.state_dict().items():
        assert np.array_equal(v.numpy(), sd[k])


def test_constructor_surface_matches_reference():
    """Argument names and defaults of the mirrored constructors."""
    import inspect
    sig = inspect.signature(TemporalUnet.__init__).parameters
    assert [p for p in sig][1:6] == ["transition_dim", "dim", "dim_mults", "kernel_size", "time_dim"]
    assert sig["dim"].default == 128 and sig["dim_mults"].default == (1, 2, 4, 8) and sig["kernel_size"].default == 5
    sig = inspect.signature(GaussianDiffusion.__init__).parameters
    assert [p for p in sig][1:] == ["model", "horizon", "observation_dim", "action_dim", "n_timesteps", "loss_type",
                                   "clip_denoised", "predict_epsilon", "beta_schedule"]
    assert sig["n_timesteps"].default == 1000 and sig["beta_schedule"].default == "cosine"
    sig = inspect.signature(DynamicsAwarePolicy.__init__).parameters
    assert [p for p in sig][1:] == ["diffusion_model", "projection_matrix", "normalizer", "state_dim", "observation_dim",
                                   "action_dim", "horizon", "projection_schedule", "projection_strength", "action_horizon"]
    sig = inspect.signature(GuidedPolicy.__init__).parameters
    assert [p for p in sig][1:] == ["diffusion_model", "normalizer", "guide_fn", "guide_weight", "action_horizon"]
    assert inspect.signature(MPCPolicy.__init__).parameters["action_horizon"].default == 8
    with pytest.raises(ValueError):
        GaussianDiffusion(TemporalUnet(6, dim=64, dim_mults=(1, 2)), 16, 4, 2, beta_schedule="sigmoid")


def test_policy_host_glue():
    """_process_observation / _fill_action_buffer / action buffering (policies.py:151-223), no device work."""
    c = helpers.CASES["tiny"]
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2))
    dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=20)
    nz = helpers.normalizer(c)
    pol = MPCPolicy(dif, nz, action_horizon=3)
    o = pol._process_observation({"observation": np.arange(4.0), "desired_goal": np.ones(2), "achieved_goal": np.zeros(2)})
    assert o.shape == (1, 4) and np.array_equal(o[0], np.arange(4.0))
    assert pol._process_observation({"achieved_goal": np.ones(3)}).shape == (1, 3)
    assert pol._process_observation(np.arange(4.0)).shape == (1, 4)
    plan = torch.zeros(1, 16, 6)
    plan[0, :, 4:] = torch.arange(32.0).reshape(16, 2)
    pol._fill_action_buffer(plan)
    assert len(pol.action_buffer) == 4            # range(0, action_horizon + 1)   (policies.py:188)
    np.testing.assert_allclose(pol.action_buffer[1], nz.unnormalize_actions(np.array([[2.0, 3.0]]))[0], rtol=1e-6)
    first = pol.get_action(np.zeros(4))           # pops the buffer: no replanning, no device needed
    np.testing.assert_allclose(first, nz.unnormalize_actions(np.array([[0.0, 1.0]]))[0], rtol=1e-6)
    assert len(pol.action_buffer) == 3


def test_no_cpu_fallback():
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2))
    with pytest.raises(RuntimeError, match="no CPU"):
        net(torch.zeros(2, 16, 6), torch.zeros(2, dtype=torch.long))
    dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=20)
    with pytest.raises(RuntimeError):
        dif.p_sample_loop((2, 16, 6))
    with pytest.raises(NotImplementedError):
        dif.loss(torch.zeros(2, 16, 6))


def test_c_abi_exports_every_declared_symbol(built_lib):
    """The shared library loads and exports exactly what include/dad_b200.h declares (no compute calls)."""
    hdr = open(os.path.join(helpers.ROOT, "include", "dad_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char \*)\s*\*?\s*(dad_[a-z_0-9]+)\s*\(", hdr, flags=re.M))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dad_abi_version() == _native.DAD_ABI_VERSION


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")
def test_create_fails_loudly_without_a_gpu(built_lib):
    from dynamics_aware_diffusion_b200.engine import Engine
    with pytest.raises(_native.DadError) as e:
        Engine(transition_dim=6, dim=64, dim_mults=(1, 2), kernel_size=5, time_dim=None, horizon=16, n_timesteps=10,
               precision="fp32", max_batch=4, device="cuda:0")
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(helpers.ROOT, "dynamics_aware_diffusion_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


@pytest.mark.parametrize("name", ["pointmaze", "cheetah_s", "door_s"])
def test_checkpoint_arch_inference(name, tmp_path):
    """infer_model_config_from_checkpoint (evaluate.py:64-122) incl. the (1,4,8) case the reference mis-infers, and
    load_diffusion round-tripping a training checkpoint (utils/training.py:193-211) with optional EMA weights."""
    from dynamics_aware_diffusion_b200 import infer_model_config_from_checkpoint, load_diffusion
    c = helpers.CASES[name]
    sd, dif = helpers.make_state_dict(c)
    ema = {k: v.clone() + 0.5 for k, v in dif.named_parameters()}
    ckpt = {"epoch": 1, "global_step": 5, "model_state_dict": dif.state_dict(), "ema_state_dict": ema,
            "config": {"horizon": c["H"], "observation_dim": c["n"], "action_dim": c["m"], "n_timesteps": c["S"],
                       "beta_schedule": c["beta"]}}
    cfg = infer_model_config_from_checkpoint(ckpt)
    assert cfg["dim"] == c["dim"] and tuple(cfg["dim_mults"]) == tuple(c["mults"]) and cfg["n_timesteps"] == c["S"]
    assert cfg["horizon"] == c["H"] and cfg["beta_schedule"] == c["beta"] and cfg["transition_dim"] == c["n"] + c["m"]
    path = tmp_path / "checkpoint_best.pt"
    torch.save(ckpt, path)
    model, _ = load_diffusion(str(path), device="cpu")
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), sd[k]), k
    model_ema, _ = load_diffusion(ckpt, device="cpu", use_ema=True)
    k0 = "model.time_mlp.1.weight"
    assert np.allclose(model_ema.state_dict()[k0].numpy(), sd[k0] + 0.5)
    with pytest.raises(ValueError):
        load_diffusion(ckpt, observation_dim=c["n"] + 1, action_dim=c["m"], device="cpu")


@pytest.mark.parametrize("name", CASE_NAMES)
def test_dynamics_residual_matches_reference(name):
    """dynamics_residual == ProjectionLoss.compute on the reference's own trajectories (golden residual_dyn)."""
    from dynamics_aware_diffusion_b200 import dynamics_residual
    c = helpers.CASES[name]
    g = helpers.load_golden(name)
    P = g["P"] if "P" in g else ProjectionMatrixBuilder(g["A"], g["Bm"], c["n"], c["m"]).get_projection_matrix(c["H"]).numpy()
    nz = helpers.normalizer(c)
    for k in (0, c["S"] - 1):
        got = dynamics_residual(g["trace_dyn"][k], P, nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std, c["n"], c["m"])
        want = float(g["residual_dyn"][k])
        assert abs(got - want) <= 1e-4 * max(want, 1e-3)


def test_optional_knobs_do_not_change_the_reference_surface():
    """The extras of this round (latency kernels, captured guidance, device projector build) are optional and
    host-visible without a GPU: defaults keep the reference's behaviour, misuse raises."""
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2), latency_max_batch=0)
    assert net.latency_max_batch == 0 and TemporalUnet(6, dim=64, dim_mults=(1, 2)).latency_max_batch is None
    dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=20)
    pol = GuidedPolicy(dif, helpers.normalizer(helpers.CASES["tiny"]), guide_fn=lambda x, t: x.sum((1, 2)))
    assert pol.capture_guidance and pol._guided_graphs == {}
    A, B = helpers.dynamics(helpers.CASES["tiny"])
    b = ProjectionMatrixBuilder(A, B, 4, 2)
    P = b.get_projection_matrix(16)                         # default: the reference's numpy path, CPU fp32 tensor
    assert P.dtype == torch.float32 and P.device.type == "cpu" and b.verify_projection(P)
    with pytest.raises(ValueError):
        b.get_projection_matrix(16, device="cpu")
    if not torch.cuda.is_available():
        # the device build needs a GPU and says so through the C-ABI error channel
        with pytest.raises(_native.DadError):
            b.get_projection_matrix(16, device="cuda:0")


def test_ill_conditioned_step_detection_is_index_based():
    """GaussianDiffusion.ill_conditioned_min_step: the trailing run of step indices with d(mean)/d(eps) > 10 -- index S-1
    only for the cosine schedule (beta clipped to 0.9999, diffusion.py:41), none for the linear one; a shortened loop
    (n_timesteps lowered after construction, evaluate.py:351-353) never reaches it."""
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion
    net = TemporalUnet(6, dim=32, dim_mults=(1, 2))
    cos = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=100, beta_schedule="cosine")
    assert cos.fp32_ill_conditioned_steps is True and cos.ill_conditioned_math == "tf32"
    assert cos.ill_conditioned_min_step() == 99 and cos.ill_conditioned_prefix() == 1
    assert cos.ill_conditioned_prefix(50) == 0
    amp = (cos.posterior_mean_coef1 * cos.sqrt_recipm1_alphas_cumprod)
    assert float(amp[99]) > 10.0 and float(amp[:99].max()) < 1.5
    lin = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=100, beta_schedule="linear")
    assert lin.ill_conditioned_min_step() == 100 and lin.ill_conditioned_prefix() == 0


def test_bench_arms_share_one_config_object():
    """bench.py: the b200 arm and the reference arm print the SAME `config` (workload_config), for every N and scaling."""
    import bench
    w = bench.WORKLOADS["pointmaze"]
    for world, scaling in ((1, "strong"), (8, "strong"), (8, "weak")):
        c = bench.workload_config("pointmaze", w, world, scaling)
        assert c["B_total"] == (4096 if scaling == "strong" else 4096 * world) and c["B_per_gpu"] * world == c["B_total"]
        assert c["diffusion_steps"] == 500 and c["H"] == 32 and c["T"] == 6 and c["policy"] == "dynamics-aware"
        assert "model" not in c
    assert bench.workload_config("pointmaze_guided", bench.WORKLOADS["pointmaze_guided"], 1, "strong")["policy"] == "guided"


def test_launch_share_tool_separates_the_sibling_step(tmp_path):
    """tools/launch_shares.py: one regular diffusion step (conv_chain launches) and the ill-conditioned step of the fp32
    sibling are reported separately from an ncu launch list."""
    import subprocess
    import sys as _sys
    rows = ['"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"']
    seq = (["dad::stage_x_kernel(LoopState *)", "void dad::conv_tf32_kernel<0>(ConvF32Params)", "dad::gn_mish_f32_kernel(GnF32Params)",
            "void dad::step_project_fused_kernel<7, 1>(StepParams)"] +
           ["dad::stage_x_kernel(LoopState *)", "void dad::conv_chain_kernel<64, 1, 2>(ChainArgs)",
            "void dad::conv_tc_kernel<128, 0>(CUtensorMap_st)", "void dad::step_project_fused_kernel<7, 1>(StepParams)"] * 2)
    for i, name in enumerate(seq):
        rows.append('"%d","1","python","box","%s","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","%d"'
                    % (i, name, 1000 * (i + 1)))
    src = tmp_path / "launches.csv"
    src.write_text("\n".join(rows) + "\n")
    dst = tmp_path / "shares.md"
    out = subprocess.run([_sys.executable, os.path.join(helpers.ROOT, "tools", "launch_shares.py"), str(src), str(dst), "cmd"],
                         capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    text = dst.read_text()
    assert "regular diffusion step" in text and "conv_chain_kernel<64, 1, 2>" in text
    assert "ill-conditioned leading step" in text and "conv_tf32_kernel<0>" in text
    assert text.index("conv_chain_kernel") < text.index("conv_tf32_kernel")
