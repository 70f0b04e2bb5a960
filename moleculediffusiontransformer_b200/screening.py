"""Closed-loop screening: inverse model -> tokens -> forward model, without leaving the device.

The reference does this per molecule through text (generative.py:1249-1261 -> predict_properties_from_SMILES,
generative.py:664-711): argmax tokens -> Keras ``sequences_to_texts`` (drops padding id 0) -> re-tokenise -> divide by
``X_norm_factor`` -> zero-pad to the forward model's context length -> ``QMDiffusionForward.sample``.  Token ids survive the
text round trip unchanged, so the same conditioning is built here directly from the uint8 tokens the inverse plan emits.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
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
    _, tokens = inverse.sample(sequences, device, cond_scale=cond_scale, timesteps=timesteps, seed=seed,
                               precision=precision, return_tokens=True)
    ctx = forward.unet.fixed_embedding.max_length
    cond = tokens_to_forward_conditioning(tokens, ctx, x_norm_factor)
    pred = forward.sample(cond, device, cond_scale=1.0, timesteps=forward_timesteps or timesteps,
                          seed=None if seed is None else seed + 1, precision=precision)
    return tokens, pred


def vocabulary_table(index_word: Dict[int, str]) -> np.ndarray:
    """Keras ``Tokenizer.index_word`` (token id -> character) -> the 256-entry byte table ``mdt_op_decode_tokens`` reads.

    The SMILES tokenisers of the reference are character level (``Tokenizer(char_level=True)``, notebooks), so every entry is one
    ASCII character; anything else is rejected rather than silently mangled.  Ids without an entry (0 = padding always) map to 0
    and are dropped by the decoder, which is what ``sequences_to_texts`` does with unknown ids when no ``oov_token`` is set."""
    lut = np.zeros(256, dtype=np.uint8)
    for idx, word in index_word.items():
        idx = int(idx)
        if not 0 < idx < 256:
            raise ValueError(f"token id {idx} outside 1..255")
        if len(word) != 1 or ord(word) >= 128 or word in (" ", "\0"):
            raise ValueError(f"token {idx} -> {word!r}: the device decoder needs single non-space ASCII characters")
        lut[idx] = ord(word)
    return lut


def reverse_tokenize(index_word: Dict[int, str], tokens: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> List[str]:
    """Device-side ``reverse_tokenize`` (generative.py:1069-1078): uint8 ``[B, L]`` argmax tokens on a CUDA device (what
    ``sample(..., return_tokens=True)`` returns) -> list of ``B`` strings.  The vocabulary lookup and the removal of padding run in
    one kernel; only ``B x L`` text bytes and ``B`` lengths cross to the host."""
    from . import _capi

    if tokens.dim() != 2 or tokens.dtype != torch.uint8 or not tokens.is_cuda:
        raise ValueError("tokens must be a uint8 [B, L] CUDA tensor")
    tokens = tokens.contiguous()
    b, l = tokens.shape
    if b == 0:
        return []
    lib = _capi.load()
    lut = torch.from_numpy(vocabulary_table(index_word)).to(tokens.device)
    out = torch.empty((b, l), dtype=torch.uint8, device=tokens.device)
    lengths = torch.empty((b,), dtype=torch.int32, device=tokens.device)
    s = stream or torch.cuda.current_stream(tokens.device)
    with torch.cuda.device(tokens.device):
        _capi.check(lib.mdt_op_decode_tokens(tokens.data_ptr(), lut.data_ptr(), out.data_ptr(), lengths.data_ptr(), b, l, s.cuda_stream))
    text = out.cpu().numpy()
    n = lengths.cpu().numpy()
    return [text[i, : n[i]].tobytes().decode("ascii") for i in range(b)]


def is_novel(all_smiles: Iterable[str], smi: str) -> bool:
    """generative.py:1063-1067."""
    return smi not in all_smiles
