"""Closed-loop screening: inverse model -> tokens -> forward model, without leaving the device.

The reference does this per molecule through text (generative.py:1249-1261 -> predict_properties_from_SMILES,
generative.py:664-711): argmax tokens -> Keras ``sequences_to_texts`` (drops padding id 0) -> re-tokenise -> divide by
``X_norm_factor`` -> zero-pad to the forward model's context length -> ``QMDiffusionForward.sample``.  Token ids survive the
text round trip unchanged, so the same conditioning is built here directly from the uint8 tokens the inverse plan emits.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def tokens_to_forward_conditioning(tokens: torch.Tensor, context_length: int, x_norm_factor: float) -> torch.Tensor:
    """uint8 [B, L] token ids (0 = padding) -> float32 [B, context_length] conditioning of the forward model.

    Non-padding ids are compacted to the front in order (what decoding to a string and re-tokenising does), divided by
    ``x_norm_factor`` (generative.py:682-685) and zero padded / truncated to ``context_length``."""
    t = tokens.to(torch.int64)
    keep = t != 0
    # stable partition: kept positions first, original order preserved
    order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)
    compact = torch.gather(t * keep, 1, order)
    b, l = compact.shape
    out = torch.zeros((b, context_length), dtype=torch.float32, device=tokens.device)
    n = min(l, context_length)
    out[:, :n] = compact[:, :n].to(torch.float32) / float(x_norm_factor)
    return out


def generate_and_score(inverse, forward, sequences: torch.Tensor, device, *, cond_scale: float = 7.5, timesteps: int = 64,
                       forward_timesteps: Optional[int] = None, x_norm_factor: float = 21.0, seed: Optional[int] = None,
                       precision: Optional[str] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Generate molecules for ``sequences`` with the inverse model and re-predict their properties with the forward model.

    Returns ``(tokens uint8 [B, L], predicted float32 [B, pred_dim_fwd, max_length_fwd])``; everything stays on ``device``."""
    _, tokens = inverse.sample(sequences, device, cond_scale=cond_scale, timesteps=timesteps, seed=seed if seed is not None else 0,
                               precision=precision, return_tokens=True)
    ctx = forward.unet.fixed_embedding.max_length
    cond = tokens_to_forward_conditioning(tokens, ctx, x_norm_factor)
    pred = forward.sample(cond, device, cond_scale=1.0, timesteps=forward_timesteps or timesteps,
                          seed=(seed if seed is not None else 0) + 1, precision=precision)
    return tokens, pred
