"""The conv-chain kernel (csrc/conv_chain.cuh): whole runs of ResidualTemporalBlocks (temporal_unet.py:106-122,
214-237) in one persistent launch, their convolutions synchronised through per-sample-tile counters instead of kernel
boundaries.  The arithmetic of an output element is the same at every fusion level (same K order, same epilogue), so
the levels must agree BIT FOR BIT -- any difference is a missed dependency (a tile read before it was complete)."""
import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def pointmaze():
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic
    net = TemporalUnet(6, dim=128, dim_mults=(1, 2, 4), precision="bf16", max_batch=4096, latency_max_batch=0)
    dif = GaussianDiffusion(net, horizon=32, observation_dim=4, action_dim=2, n_timesteps=40)
    synthetic.fill_state_dict(dif, 0)
    return dif.to(_dev())


def test_chain_structure(pointmaze):
    """PointMaze: 37 convolutions in 13 launches (5 chains of 5 / 5 / 9 / 5 / 5 convs + the final block's conv; the
    strided, transposed and head convs stay single), 15 kernels per diffusion step with staging and the step kernel."""
    eng = pointmaze.engine(32, _dev())
    eng.set_fusion(3)
    units = eng.units()
    chains = [u for u in units if u["is_chain"]]
    assert [u["n_layers"] for u in chains] == [5, 5, 9, 5, 5, 1], units
    assert all(u["kernel"].startswith("conv_chain_kernel") for u in chains)
    assert sum(u["n_layers"] for u in units) == len(eng.layers()) == 37
    assert eng.info()["launches_per_step"] == 15
    eng.set_fusion(2)
    assert [u["n_layers"] for u in eng.units() if u["is_chain"]] == [3, 2, 3, 2, 3, 2, 2, 2, 3, 2, 3, 2, 1]
    eng.set_fusion(0)
    assert not any(u["is_chain"] for u in eng.units()) and eng.info()["launches_per_step"] == 39
    eng.set_fusion(3)


@pytest.mark.parametrize("B", [1, 40, 600, 4096])
def test_fusion_levels_bit_identical(pointmaze, B):
    """U-Net output at fusion levels 0 (per-layer conv_t3 kernels of round 1), 1, 2, 3: identical bits, run to run too.
    B = 40 / 600 leave ragged tiles and an odd tile count (a CTA pair with one empty half); B = 1 is a single tile."""
    eng = pointmaze.engine(32, _dev())
    g = torch.Generator(device=_dev()).manual_seed(100 + B)
    x = torch.randn(B, 32, 6, device=_dev(), generator=g)
    outs = {}
    for level in (0, 1, 2, 3):
        eng.set_fusion(level)
        outs[level] = eng.unet_forward(x, step=7)
        for _ in range(2):
            assert torch.equal(eng.unet_forward(x, step=7), outs[level]), "run-to-run variation at level %d" % level
    torch.cuda.synchronize()
    assert bool(torch.isfinite(outs[0]).all())
    for level in (1, 2, 3):
        assert torch.equal(outs[level], outs[0]), "fusion level %d differs from the per-layer kernels" % level


def test_sampling_loop_fusion_levels_equal(pointmaze):
    """The whole captured loop (graph replays, Philox noise, projector, inpainting): chains vs per-layer kernels."""
    from dynamics_aware_diffusion_b200 import DynamicsAwarePolicy, ProjectionMatrixBuilder, synthetic
    A, Bm = synthetic.double_integrator(0.1)
    P = ProjectionMatrixBuilder(A, Bm, 4, 2).get_projection_matrix(32)
    pol = DynamicsAwarePolicy(pointmaze, projection_matrix=P, normalizer=synthetic.SyntheticNormalizer(4, 2), state_dim=4,
                              observation_dim=4, action_dim=2, horizon=32, projection_schedule="noise_schedule")
    eng = pol._engine(_dev())
    start = torch.zeros(1, 6, device=_dev())
    start[0, :4] = torch.tensor([0.3, -0.2, 0.1, 0.0])
    outs = {}
    for level in (0, 3):
        eng.set_fusion(level)
        torch.manual_seed(1)
        outs[level] = pol.sample_loop(batch_size=300, conditions={0: start}, seed=5)
    eng.set_fusion(3)
    assert bool(torch.isfinite(outs[3]).all())
    assert torch.equal(outs[0], outs[3])


def test_time_unit_matches_layers(pointmaze):
    """The measurement hook for whole launch units runs (its chain counters count epochs instead of being reset)."""
    eng = pointmaze.engine(32, _dev())
    eng.set_fusion(3)
    units = eng.units()
    big = max((u for u in units if u["is_chain"]), key=lambda u: u["flops_per_sample"])
    ms = eng.time_unit(big["index"], 512, iters=5)
    assert 0.0 < ms < 50.0
    # the counters are back at zero: a forward after the timing loop is still exact
    x = torch.randn(64, 32, 6, device=_dev(), generator=torch.Generator(device=_dev()).manual_seed(1))
    a = eng.unet_forward(x, step=3)
    eng.set_fusion(0)
    b = eng.unet_forward(x, step=3)
    eng.set_fusion(3)
    assert torch.equal(a, b)
