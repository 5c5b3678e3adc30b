"""CPU oracle for the m_diffuser reverse-diffusion sampling path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this package, and
only as the checker (or the timed CPU baseline).  The shipped path
(`dynamics_aware_diffusion_b200`) never imports it and has no CPU fallback.

What it restates (numpy, fp32 or fp64 selectable), each function citing the
reference file:line it follows:

  oracle.unet        TemporalUnet.forward          m_diffuser/models/temporal_unet.py
  oracle.diffusion   schedules, p_mean_variance,   m_diffuser/models/diffusion.py
                     p_sample, guided step,        m_diffuser/guides/policies.py:48-149
                     conditions, sampling loops
  oracle.projection  F / P = F F^+ builder,        m_diffuser/dynamics/projection.py
                     apply_projection (literal),   m_diffuser/guides/policies.py:358-485
                     fit_linear_dynamics,          m_diffuser/dynamics/data_driven.py:75-134
                     dynamics residual             m_diffuser/losses/__init__.py:161-186

The arithmetic of the reference lives in PyTorch (torch>=2.0.0, requirements.txt:2;
this container: torch 2.11.0+cu128 / oneDNN on CPU), a third-party dependency that
is not vendored in the reference tree.  The oracle restates the published
definitions of Conv1d / ConvTranspose1d / GroupNorm / Mish / Linear.

Parity pinning: the reference holds NO golden vectors for this path (SURVEY.md
F8).  The oracle is pinned instead against outputs of the reference itself, run
in the build container through `oracle/ref_shim.py`; those outputs are committed
as `tests/golden/*.npz` together with the generating script
`tests/golden/make_golden.py`, and `tests/test_oracle_golden.py` checks the
oracle against every one of them.
"""
