"""Import shim for the UNMODIFIED reference (test infrastructure only).

The reference's package __init__ imports a module that is not in its tree
(m_diffuser/__init__.py:12 -> m_diffuser.datasets) and its dynamics package
pulls gymnasium/minari (m_diffuser/dynamics/__init__.py:2-4).  We pre-seed empty
package objects whose __path__ points into /root/reference so that the hot-path
modules import normally.  Users: `tests/golden/make_golden.py`, the optional
`-m "not gpu"` live-reference tests, and bench.py's reference arm / cpu_baseline
leg.  Root: $DAD_REFERENCE_ROOT, else /root/reference (this container), else
oracle/_ref (the byte-for-byte staging of the path's files made by
oracle/make_ref.py, git-ignored, which is what exists on the GPU box).  Nothing
in the product imports it.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _root():
    env = os.environ.get("DAD_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isdir(os.path.join(cand, "m_diffuser", "models")):
            return cand
    return "/root/reference"


REF_ROOT = _root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "m_diffuser", "models"))


def load():
    """Returns a namespace with the reference classes of the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name, sub in (("m_diffuser", "m_diffuser"), ("m_diffuser.dynamics", "m_diffuser/dynamics")):
        if name not in sys.modules or not getattr(sys.modules[name], "__path__", None):
            mod = types.ModuleType(name)
            mod.__path__ = [os.path.join(REF_ROOT, sub)]
            sys.modules[name] = mod
    if "minari" not in sys.modules:
        sys.modules["minari"] = types.ModuleType("minari")
    from m_diffuser.models.temporal_unet import TemporalUnet
    from m_diffuser.models.diffusion import GaussianDiffusion, cosine_beta_schedule, linear_beta_schedule
    from m_diffuser.guides.policies import GuidedPolicy, MPCPolicy, ValueGuidedPolicy, DynamicsAwarePolicy
    from m_diffuser.dynamics.projection import ProjectionMatrixBuilder
    from m_diffuser.dynamics.data_driven import fit_linear_dynamics
    ns = types.SimpleNamespace(**{k: v for k, v in locals().items() if k[0].isupper() or k.endswith("_schedule") or k.startswith("fit_")})
    return ns
