"""Seeded parity cases shared by oracle/make_golden.py and tests/ (inputs are regenerated from seeds)."""
from __future__ import annotations

import torch

INV64 = dict(max_length=64, pred_dim=16, channels=64, unet_type="cfg", context_embedding_max_length=12,
             pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)
FWD64 = dict(max_length=64, pred_dim=1, channels=64, unet_type="cfg", context_embedding_max_length=64,
             pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)
WIDE = dict(max_length=128, pred_dim=32, channels=128, unet_type="cfg", context_embedding_max_length=12,
            pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)
PAPER = dict(max_length=32, pred_dim=22, channels=128, unet_type="cfg", context_embedding_max_length=12,
             pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)

# sibling wrappers / UNet types on the same executor: unet_type='base' (generative.py:786-810, 93-117; patch_size 8, so
# lengths >= 128) and the graphmodel.py wrappers AnalogDiffusionSparse (patch 8) / AnalogDiffusionFull (patch 4)
BASE_INV = dict(INV64, max_length=128, unet_type="base")
BASE_FWD = dict(FWD64, max_length=128, unet_type="base")
ANALOG_SPARSE = dict(max_length=128, pred_dim=3, channels=64, unet_type="cfg", context_embedding_max_length=12,
                     pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)
ANALOG_FULL = dict(max_length=64, pred_dim=8, channels=64, unet_type="cfg", context_embedding_max_length=12,
                   pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)

# name -> (kind, ctor kwargs, model seed, data seed, batch, ctx_len, cond_scale, timesteps, clamp)
CASES = {
    "inv64_cs1":    ("inverse", INV64, 0, 1234, 4, 12, 1.0, 64, False),
    "inv64_cs7p5":  ("inverse", INV64, 0, 1234, 4, 12, 7.5, 64, False),
    "inv64_short_ctx_clamp": ("inverse", INV64, 0, 77, 3, 7, 2.0, 8, True),
    "fwd64_cs1":    ("forward", FWD64, 0, 2, 4, 64, 1.0, 16, False),
    "fwd64_cs2":    ("forward", FWD64, 0, 2, 2, 64, 2.0, 6, False),
    "wide_cs7p5":   ("inverse", WIDE, 0, 3, 2, 12, 7.5, 6, False),
    "paper_cs2":    ("inverse", PAPER, 0, 5, 2, 12, 2.0, 6, False),
    # BASELINE.json configs[3] at its real depth: 128 timesteps = 254 denoiser calls x 2 branches for rounding error to accumulate over
    "wide_cs7p5_t128": ("inverse", WIDE, 0, 31, 4, 12, 7.5, 128, False),
    "base128_inv":  ("inverse", BASE_INV, 0, 51, 3, 12, 7.5, 6, False),        # cond_scale is ignored by the 'base' branch
    "base128_fwd":  ("forward", BASE_FWD, 0, 52, 2, 64, 1.0, 6, True),
    "analog_sparse": ("analog_sparse", ANALOG_SPARSE, 0, 53, 3, 12, 2.0, 6, False),
    "analog_full":  ("analog_full", ANALOG_FULL, 0, 54, 2, 9, 7.5, 6, False),
}


def make_inputs(name: str):
    kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
    g = torch.Generator().manual_seed(dseed)
    if kind == "forward":
        # synthetic tokenised SMILES / 21 (generative.py:682-685): ragged lengths, zero padded
        lens = torch.randint(1, 30, (b,), generator=g)
        toks = torch.randint(1, 22, (b, n), generator=g).float()
        toks = toks * (torch.arange(n)[None, :] < lens[:, None])
        seq = toks / 21.0
    else:
        seq = torch.rand(b, n, generator=g) * 2 - 1
    p, l = kw["pred_dim"], kw["max_length"]
    noise0 = torch.randn(b, p, l, generator=g)
    step_noise = torch.randn(steps - 1, b, p, l, generator=g)
    return seq, noise0, step_noise


# inpainting cases: name -> (ctor kwargs, model seed, data seed, batch, ctx_len, cond_scale, timesteps, num_resamples, keep_positions)
INPAINT_CASES = {
    "inpaint_inv64_r1": (INV64, 0, 41, 3, 12, 2.0, 6, 1, 20),
    "inpaint_inv64_r2": (INV64, 0, 42, 2, 12, 7.5, 5, 2, 33),
}


def make_inpaint_inputs(name: str):
    kw, mseed, dseed, b, n, cs, steps, resamples, keep = INPAINT_CASES[name]
    g = torch.Generator().manual_seed(dseed)
    seq = torch.rand(b, n, generator=g) * 2 - 1
    p, l = kw["pred_dim"], kw["max_length"]
    # a one-hot-like draft in [-1, 1] (what inpaint_from_draft_and_conditioning builds, generative.py:1574-1660)
    tok = torch.randint(0, p, (b, l), generator=g)
    source = torch.nn.functional.one_hot(tok, p).permute(0, 2, 1).float() * 2 - 1
    mask = torch.zeros(b, p, l, dtype=torch.bool)
    mask[:, :, :keep] = True                      # keep the first `keep` positions (sequential_mask, diffusion.py:628-632)
    draws = torch.randn(1 + (steps - 1) * 2 * resamples, b, p, l, generator=g)
    return seq, source, mask, draws
