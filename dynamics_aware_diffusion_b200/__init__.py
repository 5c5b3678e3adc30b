"""B200-native (sm_100a) implementation of m_diffuser's reverse-diffusion sampling path.

Same Python surface as the reference (TemporalUnet, GaussianDiffusion, the guided / MPC / value-guided /
dynamics-aware policies, ProjectionMatrixBuilder, checkpoint state_dict layout); the device work is
hand-written CUDA behind the C ABI of include/dad_b200.h.  No CPU fallback.
"""
from .temporal_unet import TemporalUnet
from .diffusion import GaussianDiffusion, cosine_beta_schedule, linear_beta_schedule
from .policies import GuidedPolicy, MPCPolicy, ValueGuidedPolicy, DynamicsAwarePolicy
from .projection import (ProjectionMatrixBuilder, fit_linear_dynamics, fold_projection, projection_alphas,
                         dynamics_residual)
from .checkpoint import infer_model_config_from_checkpoint, load_diffusion

POLICY_TYPES = ("guided", "mpc", "dynamics-aware")      # scripts/evaluate.py:38-40

__all__ = ["TemporalUnet", "GaussianDiffusion", "GuidedPolicy", "MPCPolicy", "ValueGuidedPolicy",
           "DynamicsAwarePolicy", "ProjectionMatrixBuilder", "fit_linear_dynamics", "fold_projection",
           "projection_alphas", "dynamics_residual", "infer_model_config_from_checkpoint", "load_diffusion",
           "cosine_beta_schedule", "linear_beta_schedule", "POLICY_TYPES"]
