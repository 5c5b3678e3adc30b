"""Size-independent properties at BASELINE.json's full sizes and edge cases (GPU)."""
import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def pointmaze_full():
    """PointMaze config 2: dim=128 (1,2,4), H=32, T=6, B=4096, dynamics-aware with the known double integrator."""
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, DynamicsAwarePolicy, ProjectionMatrixBuilder
    from dynamics_aware_diffusion_b200 import synthetic
    S = 500
    # latency_max_batch=0: every batch size runs the throughput kernels (the row-independence test below compares
    # small sub-batches with the full batch bit for bit)
    net = TemporalUnet(6, dim=128, dim_mults=(1, 2, 4), precision="bf16", max_batch=4096, latency_max_batch=0)
    dif = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=S)
    synthetic.fill_state_dict(dif, 0)
    dif.to(_dev())
    A, B = synthetic.double_integrator(0.1)
    P = ProjectionMatrixBuilder(A, B, 4, 2).get_projection_matrix(32)
    nz = synthetic.SyntheticNormalizer(4, 2)
    pol = DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=4, observation_dim=4, action_dim=2,
                              horizon=32, projection_schedule="noise_schedule", projection_strength=1.0)
    return dif, pol, P, nz


def test_full_size_loop_properties(pointmaze_full):
    dif, pol, P, nz = pointmaze_full
    B = 4096
    dif.n_timesteps = 12          # evaluate.py:351-353: the loop may be truncated to the first K schedule entries
    pol.n_timesteps = 12
    start = torch.zeros(1, 6, device=_dev())
    start[0, :4] = torch.tensor([0.3, -0.2, 0.1, 0.0])
    out = pol.sample_loop(batch_size=B, conditions={0: start}, seed=99)
    assert out.shape == (B, 32, 6) and bool(torch.isfinite(out).all())
    assert bool((out[:, 0] == start).all()), "inpainting must be exact"
    # determinism and shard independence: rows [1000, 1256) drawn alone with sample_offset equal the full run
    torch.manual_seed(5)
    a = pol.sample_loop(batch_size=B, conditions={0: start}, seed=99, rng="philox")
    torch.manual_seed(5)
    b = pol.sample_loop(batch_size=B, conditions={0: start}, seed=99, rng="philox")
    assert torch.equal(a, b)
    # different rows see different noise
    assert not torch.equal(a[0], a[1])
    dif.n_timesteps = 500
    pol.n_timesteps = 500


def test_projection_idempotent_and_reduces_residual(pointmaze_full):
    """P is a projector: with alpha = 1 applying the map twice equals applying it once, up to the
    extended-state trick (Q6) -- checked through the residual, which must collapse."""
    from dynamics_aware_diffusion_b200 import DynamicsAwarePolicy
    from oracle.projection import ProjectionOracle
    dif, pol, P, nz = pointmaze_full
    cpol = DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=4, observation_dim=4, action_dim=2,
                               horizon=32, projection_schedule="constant", projection_strength=1.0)
    x = torch.randn(4096, 32, 6, device=_dev(), generator=torch.Generator(device=_dev()).manual_seed(3))
    y1 = cpol.apply_projection(x, 0)
    y2 = cpol.apply_projection(y1, 0)
    orc = ProjectionOracle(P.numpy(), nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std, 4, 2, 32, 500)
    r0, r1, r2 = (orc.residual(t[:256].cpu().numpy()) for t in (x, y1, y2))
    assert r1 < 0.05 * r0 and r2 <= r1 * 1.01
    # linearity of the affine map: f(a) - f(b) = (I + N)(a - b)  =>  f(a) + f(b) - f(a + b) = f(0)
    z = torch.zeros_like(x[:8])
    fa, fb = cpol.apply_projection(x[:8], 0), cpol.apply_projection(x[8:16], 0)
    fab, f0 = cpol.apply_projection(x[:8] + x[8:16], 0), cpol.apply_projection(z, 0)
    assert helpers.rel_l2((fa + fb - fab).cpu().numpy(), f0.cpu().numpy()) < 1e-4


def test_large_batch_rows_equal_small_batch_rows(pointmaze_full):
    """Tile-scheduling regression: a sample's U-Net output must not depend on where it sits in the batch, on how
    many tiles / CTA pairs / warpgroups the launch uses, nor vary between runs (bitwise)."""
    dif, pol, P, nz = pointmaze_full
    g = torch.Generator(device=_dev()).manual_seed(11)
    x = torch.randn(4096, 32, 6, device=_dev(), generator=g)
    t = torch.full((4096,), 123, device=_dev(), dtype=torch.long)
    full = dif.model(x, t)
    again = dif.model(x, t)
    assert torch.equal(full, again), "run-to-run variation: a race in the kernels"
    for lo, n in ((0, 64), (1000, 37), (4032, 64), (2047, 2)):
        part = dif.model(x[lo:lo + n].contiguous(), t[:n])
        assert torch.equal(part, full[lo:lo + n]), "rows [%d, %d) depend on the batch they are evaluated in" % (lo, lo + n)


@pytest.mark.parametrize("arch", [dict(dim=128, mults=(1, 2, 4), H=32, T=6), dict(dim=64, mults=(1, 4, 8), H=32, T=23),
                                  dict(dim=64, mults=(1, 2, 4, 8), H=16, T=67),
                                  # the real HalfCheetah widths: 256-wide GroupNorm groups (32-channel CTAs) and
                                  # weight slabs that stream through the shared-memory ring
                                  dict(dim=256, mults=(1, 4, 8), H=32, T=23)])
def test_latency_kernels_match_throughput_kernels(arch):
    """get_action's single-plan shape: the small-batch kernels (conv_small) against the throughput kernels on the
    same weights and inputs -- same bf16 contract, different summation order -- and row independence within them."""
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic
    outs = {}
    g = torch.Generator(device=_dev()).manual_seed(3)
    x = torch.randn(8, arch["H"], arch["T"], device=_dev(), generator=g)
    t = torch.full((8,), 17, device=_dev(), dtype=torch.long)
    for mode, lat in (("throughput", 0), ("latency", 8)):
        net = TemporalUnet(arch["T"], dim=arch["dim"], dim_mults=arch["mults"], precision="bf16", max_batch=8,
                           latency_max_batch=lat)
        dif = GaussianDiffusion(net, horizon=arch["H"], observation_dim=arch["T"] - 2, action_dim=2, n_timesteps=50)
        synthetic.fill_state_dict(dif, 0)
        dif.to(_dev())
        outs[mode] = dif.model(x, t)
        if lat:
            assert torch.equal(dif.model(x, t), outs[mode]), "run-to-run variation in the latency kernels"
            for lo, n in ((0, 1), (5, 3), (7, 1)):
                part = dif.model(x[lo:lo + n].contiguous(), t[:n])
                assert torch.equal(part, outs[mode][lo:lo + n]), "latency kernels: rows depend on the batch"
            # one whole plan at B = 1, conditions in place (GuidedPolicy.sample_loop, policies.py:114-149)
            from dynamics_aware_diffusion_b200 import GuidedPolicy
            pol = GuidedPolicy(dif, synthetic.SyntheticNormalizer(arch["T"] - 2, 2))
            start = torch.zeros(1, arch["T"], device=_dev())
            start[0, 0] = 0.25
            plan = pol.sample_loop(batch_size=1, conditions={0: start}, seed=4)
            assert plan.shape == (1, arch["H"], arch["T"]) and bool(torch.isfinite(plan).all())
            assert bool((plan[:, 0] == start).all())
    err = helpers.rel_l2(outs["latency"].cpu().numpy(), outs["throughput"].cpu().numpy())
    assert err < 5e-3, err


def test_chunking_equals_single_pass():
    """B larger than the workspace capacity is processed in chunks with identical results (ragged last chunk)."""
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion
    c = helpers.CASES["tiny"]
    sd, _ = helpers.make_state_dict(c)
    outs = []
    for cap in (64, 24):
        # latency_max_batch=0: both capacities run the same (throughput) kernels, so the comparison is bitwise
        net = TemporalUnet(6, dim=64, dim_mults=(1, 2), precision="bf16", max_batch=cap, latency_max_batch=0)
        dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=20)
        dif.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        dif.to(_dev())
        g = torch.Generator(device=_dev()).manual_seed(1)
        x = torch.randn(50, 16, 6, device=_dev(), generator=g)
        z = torch.randn(20, 50, 16, 6, device=_dev(), generator=g)
        real = torch.randn
        try:
            torch.randn = lambda *a, **k: x.clone()
            outs.append(dif.p_sample_loop((50, 16, 6), noise=z))
        finally:
            torch.randn = real
    assert torch.equal(outs[0], outs[1])


def test_philox_noise_statistics_and_offsets():
    """In-kernel Philox normals: moments, determinism, and sample_offset == global row index."""
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, _native as N
    c = helpers.CASES["tiny"]
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2), precision="fp32", max_batch=8192)
    dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=20)
    dif.to(_dev())
    eng = dif.engine(16, _dev())
    B = 8192
    i = 10
    x = torch.zeros(B, 16, 6, device=_dev())
    eps = torch.zeros_like(x)
    eng.step(x, eps, i, noise=None, seed=7, sample_offset=0)
    sig = float(torch.exp(0.5 * dif.posterior_log_variance_clipped[i]))
    z = (x / sig).double()
    n = z.numel()
    assert abs(float(z.mean())) < 4 / np.sqrt(n)
    assert abs(float(z.var()) - 1) < 4 * np.sqrt(2 / n)
    assert abs(float((z ** 4).mean()) - 3) < 0.1
    x2 = torch.zeros(100, 16, 6, device=_dev())
    eng.step(x2, eps[:100].contiguous(), i, noise=None, seed=7, sample_offset=1000)
    assert torch.equal(x2, x[1000:1100])
    x3 = torch.zeros(100, 16, 6, device=_dev())
    eng.step(x3, eps[:100].contiguous(), i - 1, noise=None, seed=7, sample_offset=1000)
    assert not torch.equal(x3 / float(torch.exp(0.5 * dif.posterior_log_variance_clipped[i - 1])), x2 / sig)


def test_errors_are_loud():
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, _native as N
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2))
    with pytest.raises(RuntimeError):
        net(torch.zeros(2, 16, 6), torch.zeros(2, dtype=torch.long))        # CPU tensors: no CPU path
    dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=10).to(_dev())
    dif.n_timesteps = 11
    with pytest.raises(IndexError):
        dif.p_sample_loop((2, 16, 6))
    dif.n_timesteps = 10
    strict = TemporalUnet(6, dim=64, dim_mults=(1, 2), precision="bf16")
    with pytest.raises(N.DadError):      # 12 rows per sample do not tile the 128-row MMA: bf16 refuses, loudly
        strict(torch.zeros(2, 12, 6, device=_dev()), torch.zeros(2, dtype=torch.long, device=_dev()))


@pytest.mark.parametrize("name", ["pointmaze", "cheetah_s", "door_s"])
def test_device_projector_build_matches_numpy_pinv(name):
    """ProjectionMatrixBuilder.get_projection_matrix(H, device=cuda): fp64 Gram + Cholesky on the device against the
    reference's numpy SVD pinv (projection.py:104-107), at fp32 rounding; P is symmetric and idempotent."""
    from dynamics_aware_diffusion_b200 import ProjectionMatrixBuilder, fit_linear_dynamics
    c = helpers.CASES[name]
    dyn = helpers.dynamics(c)
    A, B = dyn if len(dyn) == 2 else fit_linear_dynamics(dyn[2], dyn[3], dyn[4])
    b = ProjectionMatrixBuilder(A, B, c["n"], c["m"])
    want = b.get_projection_matrix(c["H"]).numpy()
    got = b.get_projection_matrix(c["H"], device=_dev()).numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.abs(got - want).max() < 2e-6, np.abs(got - want).max()
    g64 = got.astype(np.float64)
    assert np.abs(g64 @ g64 - g64).max() < 1e-5 and np.abs(g64 - g64.T).max() < 1e-6


def test_device_projector_build_full_door_size_and_rank_check():
    """D = 2183, rank 935 (AdroitHand Door, H = 32): the size SURVEY.md 8(f-4) names; and a rank-deficient F is refused."""
    import time
    from dynamics_aware_diffusion_b200 import ProjectionMatrixBuilder, synthetic, _native as N
    A, B, *_ = synthetic.random_linear_system(39, 28, seed=11, n_transitions=1000)
    b = ProjectionMatrixBuilder(A, B, 39, 28)
    b.get_projection_matrix(4, device=_dev())          # context / module load
    t0 = time.perf_counter()
    got = b.get_projection_matrix(32, device=_dev()).numpy()
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = b.get_projection_matrix(32).numpy()
    t_np = time.perf_counter() - t0
    print("Door projector (2183 x 2183): device %.1f ms, numpy %.1f ms" % (t_dev * 1e3, t_np * 1e3))
    assert got.shape == (2183, 2183)
    assert np.abs(got - want).max() < 5e-6, np.abs(got - want).max()
    F = np.ones((8, 3))
    F[:, 1] = np.arange(8)
    F[:, 2] = F[:, 1] * 2 + 1          # a linear combination of the first two columns
    P = np.empty((8, 8), dtype=np.float32)
    rc = N.lib().dad_build_projection_matrix(0, F.ctypes.data, 8, 3, P.ctypes.data)
    assert rc != 0 and b"full column rank" in N.lib().dad_last_error(None)


@pytest.mark.parametrize("kind", ["mpc", "dynamics-aware", "value-guided"])
def test_get_action_end_to_end(kind):
    """GuidedPolicy.get_action (policies.py:193-223): observation -> normalise -> condition {0: [obs, 0]} -> ONE plan
    (the latency kernels' shape) -> buffered, unnormalised actions; replans only when the buffer is empty."""
    from dynamics_aware_diffusion_b200 import (TemporalUnet, GaussianDiffusion, MPCPolicy, DynamicsAwarePolicy,
                                               ValueGuidedPolicy, ProjectionMatrixBuilder, synthetic)
    net = TemporalUnet(6, dim=128, dim_mults=(1, 2, 4), precision="bf16", max_batch=8)
    dif = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=12)
    synthetic.fill_state_dict(dif, 0)
    dif.to(_dev())
    nz = synthetic.SyntheticNormalizer(4, 2)
    if kind == "mpc":
        pol = MPCPolicy(dif, nz, action_horizon=3)
    elif kind == "dynamics-aware":
        A, B = synthetic.double_integrator(0.1)
        P = ProjectionMatrixBuilder(A, B, 4, 2).get_projection_matrix(32, device=_dev())
        pol = DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=4, observation_dim=4, action_dim=2,
                                  horizon=32, projection_schedule="noise_schedule", action_horizon=3)
    else:
        value = torch.nn.Sequential(torch.nn.Linear(4, 16), torch.nn.Mish(), torch.nn.Linear(16, 1)).to(_dev())
        pol = ValueGuidedPolicy(dif, nz, value, guide_weight=0.2, action_horizon=3)
    obs = np.array([0.4, -0.1, 0.2, 0.05])
    torch.manual_seed(21)
    a0 = pol.get_action({"observation": obs, "desired_goal": np.zeros(2), "achieved_goal": obs[:2]})
    assert a0.shape == (2,) and np.isfinite(a0).all()
    assert len(pol.action_buffer) == 3                       # range(0, action_horizon + 1) filled, one popped
    # the same plan, made by hand
    torch.manual_seed(21)
    start = torch.zeros(1, 6, device=_dev())
    start[:, :4] = torch.as_tensor(nz.normalize_observations(obs[None]), dtype=torch.float32)
    plan = pol.sample_loop(batch_size=1, conditions={0: start})
    assert bool((plan[:, 0] == start).all())
    want = nz.unnormalize_actions(plan[0, :4, 4:].cpu().numpy())
    np.testing.assert_allclose(a0, want[0], rtol=1e-6, atol=1e-7)
    launches = pol._engine(_dev()).launch_count()
    for k in range(1, 4):                                     # buffered actions: no device work
        np.testing.assert_allclose(pol.get_action(obs), want[k], rtol=1e-6, atol=1e-7)
    assert pol._engine(_dev()).launch_count() == launches
    pol.get_action(obs)                                       # buffer empty -> replans
    assert pol._engine(_dev()).launch_count() > launches and len(pol.action_buffer) == 3


def test_two_policies_share_one_engine_without_stale_projector():
    """Advisor finding (round 1): the native handle is shared by every policy built on one diffusion model; the
    'projector currently loaded' tag lives on the engine, so alternating policies re-push their own N, q, alpha.  A new
    `policy.projection_matrix` / normaliser rebuilds the folded map."""
    from dynamics_aware_diffusion_b200 import (TemporalUnet, GaussianDiffusion, DynamicsAwarePolicy, ProjectionMatrixBuilder,
                                               synthetic)
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2), precision="fp32", max_batch=16)
    dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=10)
    synthetic.fill_state_dict(dif, 0)
    dif.to(_dev())
    A, B = synthetic.double_integrator(0.1)
    A2 = A.copy()
    A2[0, 2] = 0.3
    nz = synthetic.SyntheticNormalizer(4, 2)
    P1 = ProjectionMatrixBuilder(A, B, 4, 2).get_projection_matrix(16)
    P2 = ProjectionMatrixBuilder(A2, B, 4, 2).get_projection_matrix(16)
    kw = dict(normalizer=nz, state_dim=4, observation_dim=4, action_dim=2, horizon=16)
    pa = DynamicsAwarePolicy(dif, projection_matrix=P1, projection_schedule="constant", projection_strength=1.0, **kw)
    pb = DynamicsAwarePolicy(dif, projection_matrix=P2, projection_schedule="linear", projection_strength=0.5, **kw)
    assert pa._engine(_dev()) is pb._engine(_dev())

    def run(pol):
        torch.manual_seed(3)
        return pol.sample_loop(batch_size=8, seed=11)

    a1, b1 = run(pa), run(pb)
    a2, b2 = run(pa), run(pb)            # A again AFTER B pushed its projector to the shared handle
    assert torch.equal(a1, a2) and torch.equal(b1, b2)
    assert not torch.equal(a1, b1)
    # a fresh policy with A's settings gives A's result: nothing of B leaked
    pc = DynamicsAwarePolicy(dif, projection_matrix=P1, projection_schedule="constant", projection_strength=1.0, **kw)
    assert torch.equal(run(pc), a1)
    # replacing the matrix of a live policy is honoured (dynamics that change online)
    pa.projection_matrix = P2
    a3 = run(pa)
    pd = DynamicsAwarePolicy(dif, projection_matrix=P2, projection_schedule="constant", projection_strength=1.0, **kw)
    assert torch.equal(a3, run(pd)) and not torch.equal(a3, a1)
    x = torch.randn(4, 16, 6, device=_dev(), generator=torch.Generator(device=_dev()).manual_seed(1))
    assert torch.equal(pa.apply_projection(x, 0), pd.apply_projection(x, 0))


def test_weight_writes_through_data_are_seen():
    """Advisor finding (round 1): EMA.apply_shadow / update_ema / NCCL broadcasts write parameters through `.data`
    (`p.data = shadow`, `p.data.mul_()`, `p.data.copy_()`), which bumps no autograd version.  The engine must re-pack."""
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2), precision="bf16", max_batch=8)
    dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=10)
    synthetic.fill_state_dict(dif, 0)
    dif.to(_dev())
    x = torch.randn(4, 16, 6, device=_dev(), generator=torch.Generator(device=_dev()).manual_seed(1))
    t = torch.full((4,), 3, device=_dev(), dtype=torch.long)
    base = dif.model(x, t)
    p = next(net.parameters())
    # (1) in place through .data
    p.data.mul_(1.5)
    scaled = dif.model(x, t)
    assert not torch.equal(scaled, base)
    # (2) rebinding .data (EMA.apply_shadow), then restoring the original values
    shadow = p.data / 1.5
    p.data = shadow
    assert helpers.rel_l2(dif.model(x, t).cpu().numpy(), base.cpu().numpy()) < 1e-2
    # (3) the explicit hook
    with torch.no_grad():
        for q in net.parameters():
            q.data.copy_(q.data * 1.0)
    net.invalidate()
    assert bool(torch.isfinite(dif.model(x, t)).all())
    # out-of-range and non-integer timesteps are rejected instead of reading outside the tables
    with pytest.raises(IndexError):
        dif.model(x, torch.full((4,), 10, device=_dev(), dtype=torch.long))
    with pytest.raises(ValueError):
        dif.model(x, torch.full((4,), 2.5, device=_dev()))
    assert torch.equal(dif.model(x, torch.full((4,), 3.0, device=_dev())), dif.model(x, t))


def test_lean_step_kernel_equals_generic_kernel():
    """dad_sample without injected noise / trace takes the LEAN instantiations of the step kernels (loop configuration fixed
    at compile time, software-pipelined loads); with a trace it takes the generic ones.  Same Philox draws, same
    arithmetic: identical bits, for the plain policy (step_pointwise_kernel) and the dynamics-aware one (fused projector)."""
    from dynamics_aware_diffusion_b200 import (TemporalUnet, GaussianDiffusion, GuidedPolicy, DynamicsAwarePolicy,
                                               ProjectionMatrixBuilder, synthetic)
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2), precision="bf16", max_batch=512)
    dif = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=8)
    synthetic.fill_state_dict(dif, 2)
    dif.to(_dev())
    nz = synthetic.SyntheticNormalizer(4, 2)
    A, Bm = synthetic.double_integrator(0.1)
    P = ProjectionMatrixBuilder(A, Bm, 4, 2).get_projection_matrix(32)
    start = torch.zeros(1, 6, device=_dev())
    start[0, :4] = torch.tensor([0.1, 0.2, -0.3, 0.0])
    for pol in (GuidedPolicy(dif, nz),
                DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=4, observation_dim=4, action_dim=2,
                                    horizon=32, projection_schedule="noise_schedule")):
        for B in (3, 300):
            torch.manual_seed(3)
            lean = pol.sample_loop(batch_size=B, conditions={0: start}, seed=17)
            torch.manual_seed(3)
            generic, trace = pol.sample_loop(batch_size=B, conditions={0: start}, seed=17, return_trace=True)
            assert torch.equal(lean, generic), type(pol).__name__
            assert torch.equal(trace[-1], generic)


def test_tensor_core_projector_for_large_batches_matches_the_fused_kernel(pointmaze_full):
    """From B >= 8192 the D = 192 projector runs on the tensor cores (pointwise kernel + bf16x3 tcgen05 GEMM) instead of
    the fused SIMT kernel.  A plan's result does not depend on its batch (Philox subsequence = global index), so rows of a
    B = 8192 loop must equal the same rows sampled in a small batch up to the bf16x3 rounding of the projector
    (fp32-level: 3 bf16 products)."""
    dif, pol, P, nz = pointmaze_full
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, DynamicsAwarePolicy, synthetic
    net = TemporalUnet(6, dim=128, dim_mults=(1, 2, 4), precision="bf16", max_batch=8192, latency_max_batch=0)
    dif2 = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=6, beta_schedule="linear")
    synthetic.fill_state_dict(dif2, 0)
    dif2.to(_dev())
    pol2 = DynamicsAwarePolicy(dif2, projection_matrix=P, normalizer=nz, state_dim=4, observation_dim=4, action_dim=2,
                               horizon=32, projection_schedule="noise_schedule", projection_strength=1.0)
    start = torch.zeros(1, 6, device=_dev())
    start[0, :4] = torch.tensor([0.3, -0.2, 0.1, 0.0])
    # x_S from Philox too (FLAG_PHILOX_INIT: subsequence = global sample index), as bench.py's multi-GPU check does
    from dynamics_aware_diffusion_b200 import _native as N
    eng = pol2._engine(_dev())
    flags = pol2._loop_flags(eng) | N.FLAG_CONDITIONS | N.FLAG_PHILOX_INIT
    big = torch.empty(8192, 32, 6, device=_dev())
    small = torch.empty(64, 32, 6, device=_dev())
    # ONE step: identical U-Net inputs, so the only difference is the projector path (fp32-level); six free-running steps:
    # that difference passes through the bf16 rounding of the U-Net inputs and stays far below the bf16 step tolerance
    for n_steps, tol in ((1, 2e-5), (6, 5e-3)):
        eng.set_conditions({0: start}, 8192)
        eng.sample(big, n_steps, flags=flags, seed=5, sample_offset=0)
        eng.set_conditions({0: start}, 64)
        eng.sample(small, n_steps, flags=flags, seed=5, sample_offset=4000)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(big).all()) and bool((big[:, 0] == start).all())
        err = helpers.rel_l2(big[4000:4064].cpu().numpy(), small.cpu().numpy())
        assert 0 < err < tol, (n_steps, err)          # not bit-equal (another projector path)
    del pol2, dif2, net
    torch.cuda.empty_cache()


def test_fp32_sibling_api_and_math_modes():
    """dad_set_fp32_steps / dad_set_fp32_math through the raw C ABI: argument validation, and the accuracy ordering of the
    sibling's arithmetic (TF32 tensor cores < 2e-3, 3xTF32 < 5e-5, IEEE fp32 < 1e-5 on eps vs the reference)."""
    import ctypes
    from dynamics_aware_diffusion_b200 import _native as N
    import test_gpu_parity as T
    c, g, dif, sd = T.models("pointmaze", "bf16")
    eng = dif.engine(c["H"], _dev())
    comp = eng._companion
    assert comp is not None and comp[1] == c["S"] - 1 and comp[0].precision == "fp32"
    L = N.lib()
    assert L.dad_set_fp32_steps(eng.handle, eng.handle, 0) == N.ERR_INVALID          # a bf16 handle is no fp32 companion
    assert L.dad_set_fp32_steps(comp[0].handle, comp[0].handle, 0) == N.ERR_INVALID  # nor is a handle its own
    assert L.dad_set_fp32_math(eng.handle, 1) == N.ERR_INVALID                       # bf16 handles have no fp32 math mode
    assert L.dad_set_fp32_math(comp[0].handle, 7) == N.ERR_INVALID
    x = T.cu(g["x_init"])
    k = list(g["unet_steps"]).index(c["S"] - 1)
    errs = {}
    for mode in ("tf32", "tf32x3", "fp32"):
        comp[0].set_fp32_math(mode)
        errs[mode] = helpers.rel_l2(comp[0].unet_forward(x, step=c["S"] - 1).cpu().numpy(), g["unet_eps"][k])
    comp[0].set_fp32_math(dif.ill_conditioned_math)
    assert errs["tf32"] < 2e-3 and errs["tf32x3"] < 5e-5 and errs["fp32"] < 1e-5, errs
    # detaching restores the all-bf16 loop; re-attaching through engine() restores the default
    dif.fp32_ill_conditioned_steps = False
    assert dif.engine(c["H"], _dev())._companion is None
    dif.fp32_ill_conditioned_steps = True
    assert dif.engine(c["H"], _dev())._companion is not None
