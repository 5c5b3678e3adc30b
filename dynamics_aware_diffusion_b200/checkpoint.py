"""Checkpoint loading for the sampler: architecture inference from a training checkpoint and model construction.

  infer_model_config_from_checkpoint   scripts/evaluate.py:64-122 (same keys in the returned dict), with the level
                                       multipliers READ from the weights instead of guessed from the level count
                                       (the reference maps any 3-level net to (1,2,4): a HalfCheetah (1,4,8)
                                       checkpoint then fails to load)
  load_diffusion                       the model part of scripts/evaluate.py:125-203 (load_model) without the dataset:
                                       TemporalUnet + GaussianDiffusion + load_state_dict(model_state_dict | ema)

The checkpoint dict is utils/training.py:193-211: epoch, global_step, model_state_dict, optimizer_state_dict,
config{horizon, observation_dim, action_dim, n_timesteps, beta_schedule}, [ema_state_dict], [scheduler_state_dict].
"""
from typing import Optional

import torch

from .diffusion import GaussianDiffusion
from .temporal_unet import TemporalUnet


def infer_model_config_from_checkpoint(checkpoint: dict) -> dict:
    sd = checkpoint["model_state_dict"]
    saved = checkpoint.get("config", {}) or {}
    n_timesteps = sd["betas"].shape[0] if "betas" in sd else saved.get("n_timesteps", 200)
    levels = 1 + max([int(k.split(".")[2]) for k in sd if k.startswith("model.downs.") and k.split(".")[2].isdigit()],
                     default=-1)
    first = "model.downs.0.0.blocks.0.block.0.weight"
    dim = sd[first].shape[0] if first in sd else 128
    kernel_size = sd[first].shape[2] if first in sd else 5
    transition_dim = sd[first].shape[1] if first in sd else None
    if levels > 0 and first in sd:
        # width of level l = out-channels of its first conv; multiplier = width / dim (exact for reference checkpoints)
        mults = tuple(int(sd["model.downs.%d.0.blocks.0.block.0.weight" % l].shape[0]) // dim for l in range(levels))
    else:
        mults = (1, 2, 4, 8)
    return {"dim": int(dim), "dim_mults": list(mults), "n_timesteps": int(n_timesteps),
            "beta_schedule": saved.get("beta_schedule", "cosine"), "horizon": saved.get("horizon", 16),
            "kernel_size": int(kernel_size), "transition_dim": None if transition_dim is None else int(transition_dim)}


def load_diffusion(checkpoint, observation_dim: Optional[int] = None, action_dim: Optional[int] = None,
                   device="cuda", use_ema: bool = False, precision: str = "auto", max_batch: Optional[int] = None):
    """checkpoint: path or the loaded dict.  observation_dim / action_dim default to the checkpoint's config
    (the reference takes them from the dataset object, evaluate.py:182-183).  Returns (GaussianDiffusion, config)."""
    if not isinstance(checkpoint, dict):
        checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=False)      # evaluate.py:140
    cfg = infer_model_config_from_checkpoint(checkpoint)
    saved = checkpoint.get("config", {}) or {}
    observation_dim = observation_dim if observation_dim is not None else saved.get("observation_dim")
    action_dim = action_dim if action_dim is not None else saved.get("action_dim")
    if observation_dim is None or action_dim is None:
        raise ValueError("observation_dim / action_dim are neither given nor stored in the checkpoint config")
    T = observation_dim + action_dim
    if cfg["transition_dim"] is not None and cfg["transition_dim"] != T:
        raise ValueError("checkpoint was trained on transition_dim=%d, got observation_dim+action_dim=%d"
                         % (cfg["transition_dim"], T))
    net = TemporalUnet(T, dim=cfg["dim"], dim_mults=tuple(cfg["dim_mults"]), kernel_size=cfg["kernel_size"],
                       precision=precision, max_batch=max_batch)
    dif = GaussianDiffusion(net, horizon=cfg["horizon"], observation_dim=observation_dim, action_dim=action_dim,
                            n_timesteps=cfg["n_timesteps"], beta_schedule=cfg["beta_schedule"])
    state = checkpoint["model_state_dict"]
    if use_ema and "ema_state_dict" in checkpoint:
        # utils/training.py:18-62 stores the shadow parameters under their parameter names
        state = dict(state)
        state.update({k: v for k, v in checkpoint["ema_state_dict"].items() if k in state})
    dif.load_state_dict(state, strict=True)
    return dif.to(device), cfg
