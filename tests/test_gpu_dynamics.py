"""SURVEY.md 8 f-4 on the device: the data-driven dynamics fit (data_driven.py:107-121) and the dynamics-violation
metric (losses/__init__.py:161-186), both behind the C ABI, against the reference's own outputs in the golden files."""
import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("name", ["cheetah_s", "door_s"])
def test_device_fit_linear_dynamics(name):
    """dad_fit_linear_dynamics (fp64 normal equations + Cholesky) == the reference's numpy lstsq on the same data."""
    from dynamics_aware_diffusion_b200 import fit_linear_dynamics
    c, g = helpers.CASES[name], helpers.load_golden(name)
    _, _, X, U, Xn = helpers.dynamics(c)
    A, B = fit_linear_dynamics(X, U, Xn, device=_dev())
    assert A.shape == (c["n"], c["n"]) and B.shape == (c["n"], c["m"])
    np.testing.assert_allclose(A, g["A"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(B, g["Bm"], rtol=0, atol=1e-9)
    # state_dim truncation like the reference (data_driven.py:100-105) and the host path agree
    A2, B2 = fit_linear_dynamics(np.hstack([X, X[:, :2]]), U, np.hstack([Xn, Xn[:, :2]]), state_dim=c["n"], device=_dev())
    np.testing.assert_allclose(A2, A, atol=1e-12)
    np.testing.assert_allclose(B2, B, atol=1e-12)


def test_device_fit_rejects_rank_deficient_data():
    from dynamics_aware_diffusion_b200 import fit_linear_dynamics, _native as N
    rng = np.random.default_rng(0)
    X = rng.normal(size=(500, 3))
    X[:, 2] = X[:, 0]                               # duplicate column: (A, B) not identifiable
    U = rng.normal(size=(500, 2))
    with pytest.raises(N.DadError):
        fit_linear_dynamics(X, U, X, device=_dev())


@pytest.mark.parametrize("name", ["tiny", "pointmaze", "cheetah_s", "door_s"])
def test_device_dynamics_residual(name):
    """dad_dynamics_residual on the reference's traces == the residuals the reference computed for them
    (golden residual_dyn, fp32 torch) and == the numpy fp64 restatement."""
    from dynamics_aware_diffusion_b200 import dynamics_residual, ProjectionMatrixBuilder, DynamicsAwarePolicy
    c, g = helpers.CASES[name], helpers.load_golden(name)
    P = g["P"] if "P" in g else ProjectionMatrixBuilder(g["A"], g["Bm"], c["n"], c["m"]).get_projection_matrix(c["H"]).numpy()
    nz = helpers.normalizer(c)
    for k in (0, c["S"] - 1):
        x = torch.from_numpy(np.ascontiguousarray(g["trace_dyn"][k])).to(_dev())
        got = dynamics_residual(x, P, nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std, c["n"], c["m"])
        host = dynamics_residual(g["trace_dyn"][k], P, nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std, c["n"], c["m"])
        want = float(g["residual_dyn"][k])
        assert abs(got - host) <= 2e-5 * max(host, 1e-6), (k, got, host)
        assert abs(got - want) <= 5e-5 * max(want, 1e-6), (k, got, want)
    # a ragged batch (not a multiple of the 16-sample block) and the policy-level entry point
    xb = torch.randn(37, c["H"], helpers.case_T(c), device=_dev(), generator=torch.Generator(device=_dev()).manual_seed(2))
    got = dynamics_residual(xb, P, nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std, c["n"], c["m"])
    host = dynamics_residual(xb.cpu().numpy(), P, nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std, c["n"], c["m"])
    assert abs(got - host) <= 2e-5 * host


def test_policy_dynamics_residual_collapses_after_sampling():
    """End to end: the residual of dynamics-aware samples is far below that of plain samples (README.md:24-25)."""
    from dynamics_aware_diffusion_b200 import (TemporalUnet, GaussianDiffusion, DynamicsAwarePolicy, ProjectionMatrixBuilder,
                                               synthetic)
    net = TemporalUnet(6, dim=64, dim_mults=(1, 2), precision="bf16", max_batch=64)
    dif = GaussianDiffusion(net, horizon=16, observation_dim=4, action_dim=2, n_timesteps=20)
    synthetic.fill_state_dict(dif, 0)
    dif.to(_dev())
    A, B = synthetic.double_integrator(0.1)
    P = ProjectionMatrixBuilder(A, B, 4, 2).get_projection_matrix(16)
    pol = DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=synthetic.SyntheticNormalizer(4, 2), state_dim=4,
                              observation_dim=4, action_dim=2, horizon=16, projection_schedule="noise_schedule")
    x_dyn = pol.sample_loop(batch_size=32, seed=1)
    pol.project_in_loop = False
    x_plain = pol.sample_loop(batch_size=32, seed=1)
    assert pol.dynamics_residual(x_dyn) < 0.05 * pol.dynamics_residual(x_plain)
