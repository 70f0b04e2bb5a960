"""Drop-in task wrappers: ``QMDiffusion`` (inverse) and ``QMDiffusionForward`` (property predictor).

Same constructor kwargs, same ``state_dict`` keys, same ``sample(sequences, device, cond_scale,
timesteps, clamp)`` contract as the reference (generative.py:720-870, 33-180).  The body of
``sample`` hands the conditioning matrix to a CUDA plan (plan.py -> csrc/): the conditioning
encoder, the 63-iteration ADPM2 loop and the UNet all run inside the sm_100a library.  There
is no eager / CPU fallback: without a CUDA device and the built extension ``sample`` raises.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .diffusion import ADPM2Sampler, KarrasSchedule, XDiffusion_x
from .unet_params import UNetCFG1dParams, XUNet1d


class _FourierPE(nn.Module):
    """Holds the ``inv_freq`` buffer of PositionalEncoding1D (transformer.py:3444-3454)."""

    def __init__(self, channels: int):
        super().__init__()
        self.org_channels = channels
        channels = int(np.ceil(channels / 2) * 2)
        self.channels = channels
        inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
        self.register_buffer("inv_freq", inv_freq)


class _QMBase(nn.Module):
    _default_cond_scale = 1.0
    _unet_kwargs: dict = {}
    # the 'base' branch of every wrapper builds the same UNet (generative.py:790-802, 97-109; graphmodel.py:296-308, 464-476)
    _base_unet_kwargs: dict = dict(patch_size=8, multipliers=[1, 2, 4], factors=[4, 4], num_blocks=[2, 2], attentions=[1, 1],
                                   attention_heads=8, attention_features=64, attention_multiplier=2, attention_use_rel_pos=False)

    def __init__(self, max_length, channels, pred_dim, unet, context_embedding_max_length, unet_type,
                 pos_emb_fourier, pos_emb_fourier_add, text_embed_dim, embed_dim_position):
        super().__init__()
        if unet_type not in ("cfg", "base"):
            raise NotImplementedError(f"unet_type={unet_type!r}: the reference builds 'cfg' and 'base' only (generative.py:757-810)")
        self.unet_type = unet_type
        self.fc1 = nn.Linear(1, text_embed_dim)
        self.text_embed_dim = text_embed_dim
        self.embed_dim_position = embed_dim_position
        self.pos_emb_fourier = pos_emb_fourier
        self.pos_emb_fourier_add = pos_emb_fourier_add
        ctx_features = text_embed_dim
        if pos_emb_fourier:
            if not pos_emb_fourier_add:
                ctx_features = text_embed_dim + embed_dim_position
            self.p_enc_1d = _FourierPE(embed_dim_position)
        self.max_length = max_length
        self.pred_dim = pred_dim
        if unet is not None:
            if not isinstance(unet, UNetCFG1dParams):
                raise TypeError("unet= must be a moleculediffusiontransformer_b200 XUNet1d(type='cfg' | 'base', ...) instance")
            self.unet = unet
        elif unet_type == "cfg":
            self.unet = XUNet1d(type="cfg", in_channels=pred_dim, channels=channels,
                                context_embedding_features=ctx_features,
                                context_embedding_max_length=context_embedding_max_length,
                                **self._unet_kwargs)
        else:   # unconditional UNet1d (generative.py:786-802 / 93-109): the conditioning is encoded and then ignored
            self.unet = XUNet1d(type="base", in_channels=pred_dim, channels=channels, **self._base_unet_kwargs)
        self.diffusion = XDiffusion_x(type="k", net=self.unet, sigma_data=0.1, dynamic_threshold=0.0)
        object.__setattr__(self.diffusion, "_runner", self._run_sampler)
        self._plans = {}

    # ------------------------------------------------------------------ plan management
    def _plan_for(self, device: torch.device, precision: Optional[str] = None, batch: Optional[int] = None,
                  timesteps: int = 0):
        """One cached plan per (device, precision).  The workspace is sized for a power-of-two chunk that covers `batch`
        (capped at MDT_MAX_BATCH, default 32 rows per SM); a larger request rebuilds the plan, a smaller one reuses it."""
        from .plan import SamplerPlan, default_max_batch, default_precision

        precision = precision or default_precision()
        device = torch.device(device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        device = torch.device("cuda", index)                # 'cuda' and 'cuda:0' are one plan, not two
        key = (index, precision)
        plan = self._plans.get(key)
        version = self._weights_fingerprint()
        cap = default_max_batch(device)
        want = cap if batch is None else min(cap, max(8, 1 << (max(int(batch), 1) - 1).bit_length()))
        steps = max(256, 1 << (max(int(timesteps), 2) - 1).bit_length())      # FiLM tables are sized per denoiser call
        if plan is None or plan.weights_version != version or plan.max_batch < want or plan.max_timesteps < timesteps:
            if plan is not None:
                plan.close()
            plan = SamplerPlan(self, device, precision=precision, max_batch=want, max_timesteps=steps)
            plan.weights_version = version
            self._plans[key] = plan
        return plan

    def _weights_fingerprint(self):
        """Changes whenever a parameter or buffer is re-bound or written through an autograd-visible op.  Writes that bypass the
        version counter (``p.data.copy_(w)``, an EMA swap through ``.data``) are invisible to PyTorch itself: call
        ``invalidate_plans()`` after those."""
        tensors = list(self.parameters()) + list(self.buffers())
        return hash(tuple((t.data_ptr(), t._version) for t in tensors))

    def invalidate_plans(self):
        """Drop every packed-weight plan; the next ``sample`` re-packs the current ``state_dict``."""
        for p in self._plans.values():
            p.close()
        self._plans = {}

    def _apply(self, fn, *a, **k):  # .to()/.cuda()/.float() invalidate packed weights
        for p in self._plans.values():
            p.close()
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        for p in self._plans.values():
            p.close()
        self._plans = {}
        return super().load_state_dict(*a, **k)

    # ------------------------------------------------------------------ reference API
    def set_training_delegate(self, reference_model):
        """Training (``forward(sequences, output) -> loss``, generative.py:812-833 / 120-143) is not part of the accelerated
        path.  A user who trains hands over an instance of the reference's own class once; ``forward`` then loads the current
        ``state_dict`` into it (same keys) and returns its loss.  Call ``sync_from_training_delegate()`` before sampling."""
        object.__setattr__(self, "_training_delegate", reference_model)

    def sync_from_training_delegate(self):
        ref = getattr(self, "_training_delegate", None)
        if ref is None:
            raise RuntimeError("no training delegate set")
        self.load_state_dict(ref.state_dict())

    def forward(self, sequences, output):
        ref = getattr(self, "_training_delegate", None)
        if ref is None:
            raise NotImplementedError("training loss (generative.py:812-833) is outside the accelerated sampling path; "
                                      "set_training_delegate(<reference model>) routes it to the reference implementation")
        if not getattr(self, "_delegate_synced", False):
            ref.load_state_dict(self.state_dict())
            object.__setattr__(self, "_delegate_synced", True)
        return ref(sequences, output)

    def _run_sampler(self, *, noise, num_steps, sigma_schedule, sampler, clamp, embedding=None,
                     embedding_scale=1.0, sequences=None, step_noise=None, seed=None, precision=None,
                     return_tokens=False):
        # the reference contract (diffusion.py:724-741) passes the encoded conditioning as `embedding=`; `sequences=` (raw
        # properties, encoded on the device) is the wrapper's own fast path
        ctx = sequences if sequences is not None else embedding
        if ctx is None:
            raise ValueError("one of sequences= (raw conditioning) or embedding= (encoded, [B, n, F]) is required")
        if sequences is None and embedding.dim() != 3:
            raise ValueError("embedding= must be the encoded conditioning [B, n, context_embedding_features]")
        device = noise.device if noise is not None else ctx.device
        if seed is None and step_noise is None:
            # fresh ancestral noise per call like the reference's randn_like (diffusion.py:514), reproducible under torch.manual_seed
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        plan = self._plan_for(torch.device(device), precision, batch=ctx.shape[0], timesteps=num_steps)
        return plan.sample(ctx, noise0=noise, step_noise=step_noise, num_steps=num_steps,
                           sigma_schedule=sigma_schedule, sampler=sampler, clamp=clamp,
                           cond_scale=float(embedding_scale), seed=seed, return_tokens=return_tokens,
                           pre_encoded=sequences is None)

    def sample(self, sequences, device, cond_scale=None, timesteps=100, clamp=False, *,
               noise=None, step_noise=None, seed=None, precision=None, return_tokens=False):
        """Generate ``(B, pred_dim, max_length)`` float32 on ``device``.

        Positional arguments are the reference's (generative.py:834 / :146).  Keyword-only extras:
        ``noise`` / ``step_noise`` inject the initial and per-iteration noise tensors
        (``(B,P,L)`` and ``(timesteps-1,B,P,L)``) for parity runs; otherwise the initial noise is
        drawn from the CPU global generator exactly like the reference (generative.py:853) and the
        ancestral noise comes from an in-kernel Philox stream keyed by (seed, sample index, step).
        """
        if cond_scale is None:
            cond_scale = self._default_cond_scale
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("moleculediffusiontransformer_b200 runs on sm_100a CUDA devices only "
                               f"(got device={device}); there is no CPU path")
        b = sequences.shape[0]
        if self.unet_type == "cfg" and sequences.shape[1] > self.unet.fixed_embedding.max_length:
            raise AssertionError("Input sequence length must be <= max_length")  # modules.py:1194-1195
        if b == 0:  # nothing to launch; keep the reference's return contract
            empty = torch.empty((0, self.pred_dim, self.max_length), dtype=torch.float32, device=device)
            return (empty, torch.empty((0, self.max_length), dtype=torch.uint8, device=device)) if return_tokens else empty
        if noise is None and seed is None:
            noise = torch.randn(b, self.pred_dim, self.max_length)
        if noise is not None:
            noise = noise.to(device)
        return self.diffusion.sample(
            num_steps=timesteps, sampler=ADPM2Sampler(rho=1),
            sigma_schedule=KarrasSchedule(sigma_min=0.001, sigma_max=9.0, rho=3.0), clamp=clamp,
            noise=noise, embedding=None, embedding_scale=cond_scale, sequences=sequences.to(device),
            step_noise=step_noise, seed=seed, precision=precision, return_tokens=return_tokens)

    def inpaint(self, sequences, device, cond_scale=7.5, timesteps=100, num_resamples=1, inpaint=None, in_paint_mask=None, *,
                noise=None, seed=None, precision=None):
        """Conditional inpainting (generative.py:871-914 / 182-225): positions where ``in_paint_mask`` is True keep ``inpaint``.

        Positional arguments are the reference's.  ``noise`` optionally injects every RNG draw of ADPM2Sampler.inpaint
        (diffusion.py:526-549) as a ``(1 + (timesteps-1)*2*num_resamples, B, P, L)`` tensor in draw order; otherwise the draws
        come from the in-kernel Philox stream (seed taken from torch's CPU generator unless given)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("moleculediffusiontransformer_b200 runs on sm_100a CUDA devices only "
                               f"(got device={device}); there is no CPU path")
        if inpaint is None or in_paint_mask is None:
            raise ValueError("inpaint and in_paint_mask are required")
        if self.unet_type == "cfg" and sequences.shape[1] > self.unet.fixed_embedding.max_length:
            raise AssertionError("Input sequence length must be <= max_length")  # modules.py:1194-1195
        if seed is None and noise is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        plan = self._plan_for(device, precision, batch=sequences.shape[0], timesteps=timesteps)
        return plan.inpaint(sequences.to(device), inpaint, in_paint_mask, num_steps=timesteps, num_resamples=num_resamples,
                            sigma_schedule=KarrasSchedule(sigma_min=0.001, sigma_max=9.0, rho=3.0), sampler=ADPM2Sampler(rho=1),
                            cond_scale=float(cond_scale), noise=noise, seed=seed)


class QMDiffusion(_QMBase):
    """Inverse model: 12 properties -> (pred_dim, max_length) token logits (generative.py:718-914)."""

    _default_cond_scale = 7.5
    _base_unet_kwargs = dict(_QMBase._base_unet_kwargs, pre_transformer=2)      # generative.py:790
    _unet_kwargs = dict(pre_transformer=2, patch_size=1, multipliers=[1, 2, 4], factors=[4, 4],
                        num_blocks=[3, 3], attentions=[4, 4], attention_heads=8, attention_features=64,
                        attention_multiplier=2, attention_use_rel_pos=False)

    def __init__(self, max_length=1024, channels=128, pred_dim=1, context_embedding_max_length=32,
                 unet_type="cfg", pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=1024,
                 embed_dim_position=64, unet=None):
        super().__init__(max_length, channels, pred_dim, unet, context_embedding_max_length, unet_type,
                         pos_emb_fourier, pos_emb_fourier_add, text_embed_dim, embed_dim_position)


class QMDiffusionForward(_QMBase):
    """Forward model: SMILES token embedding -> property vector (generative.py:31-225)."""

    _default_cond_scale = 1.0
    _unet_kwargs = dict(patch_size=4, multipliers=[1, 2, 4], factors=[4, 4], num_blocks=[3, 3],
                        attentions=[2, 2], attention_heads=8, attention_features=64,
                        attention_multiplier=2, attention_use_rel_pos=False)

    def __init__(self, max_length=1024, channels=128, pred_dim=1, unet=None, context_embedding_max_length=32,
                 unet_type="cfg", pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=1024,
                 embed_dim_position=64):
        super().__init__(max_length, channels, pred_dim, unet, context_embedding_max_length, unet_type,
                         pos_emb_fourier, pos_emb_fourier_add, text_embed_dim, embed_dim_position)
