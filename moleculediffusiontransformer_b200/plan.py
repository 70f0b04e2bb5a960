"""SamplerPlan: Python owner of one ``mdt_plan`` (C ABI) for one model on one device.

Responsibilities kept on the host: hand the state_dict to the library by name, compute the
sigma schedule / per-iteration scalars with the reference's exact float32/double arithmetic
(diffusion.py), and pass torch-owned device buffers + the current CUDA stream across the ABI.
Everything else (conditioning encoder, UNet, sampler loop) runs inside libmdt_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _capi
from .diffusion import ITER_SCALAR_FIELDS, KarrasSampler, build_iter_scalars, karras_noise_scales


def default_precision() -> str:
    """fp32 | tf32 | bf16 | fp16.  Default: fp16 operands with fp32 accumulation -- the operands carry tf32's 11-bit significand in
    half the bytes, so the mode holds the same 1e-3 / 99.9 % parity bound as tf32 (tests/test_gpu_parity.py runs both against the
    same reference fixtures) at bf16-mode speed.  Conversions saturate at +-65504; use MDT_PRECISION=tf32 (or precision="tf32")
    for checkpoints whose normalised activations or weights could leave the fp16 range."""
    return os.environ.get("MDT_PRECISION", "fp16")


def default_max_batch(device=None) -> int:
    """Rows per chunk of a long sweep (one plan workspace, one graph replay per chunk).  MDT_MAX_BATCH if set; otherwise 32 rows per
    SM (4736 on a 148-SM B200): the level-2 GEMMs see 8 B / 128 row blocks, and with B = 32 x SMs every level of the README
    architecture is a whole number of waves (296 / 1184 / 4736 row blocks), where B = 4096 leaves the 256-block level at 1.73 waves."""
    env = os.environ.get("MDT_MAX_BATCH")
    if env:
        return int(env)
    if torch.cuda.is_available():
        return 32 * torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count
    return 4096


def make_config(model, precision: str, max_batch: int, max_timesteps: int) -> _capi.MdtConfig:
    u = model.unet.cfg
    cfg = _capi.MdtConfig()
    cfg.abi_version = _capi.MDT_ABI_VERSION
    cfg.in_channels, cfg.out_channels, cfg.length = u.in_channels, u.out_channels, model.max_length
    cfg.channels, cfg.patch_size, cfg.num_levels = u.channels, u.patch_size, u.num_levels
    if u.num_levels > _capi.MDT_MAX_LEVELS:
        raise ValueError("too many UNet levels")
    for i, m in enumerate(u.multipliers):
        cfg.multipliers[i] = m
    for i in range(u.num_levels):
        cfg.factors[i], cfg.num_blocks[i], cfg.attentions[i] = u.factors[i], u.num_blocks[i], u.attentions[i]
    cfg.attentions[u.num_levels] = u.attentions[-1]          # bottleneck depth (modules.py:1063)
    cfg.pre_transformer = u.pre_transformer
    cfg.heads, cfg.head_features, cfg.ff_multiplier = u.attention_heads, u.attention_features, u.attention_multiplier
    cfg.resnet_groups = u.resnet_groups
    cfg.kernel_multiplier_downsample = u.kernel_multiplier_downsample
    cfg.use_skip_scale = int(u.use_skip_scale)
    cfg.mapping_features = u.mapping_features
    # XUNet1d(type='base') takes no conditioning: ctx_features = 0 switches the encoder, cross-attention and guidance off
    cfg.ctx_features, cfg.ctx_max_length = u.context_embedding_features or 0, u.context_embedding_max_length or 0
    cfg.text_embed_dim, cfg.embed_dim_position = model.text_embed_dim, model.embed_dim_position
    cfg.pos_emb_fourier, cfg.pos_emb_fourier_add = int(model.pos_emb_fourier), int(model.pos_emb_fourier_add)
    cfg.sigma_data = model.diffusion.diffusion.sigma_data
    cfg.precision = _capi.PRECISIONS[precision]
    cfg.max_batch, cfg.max_timesteps = max_batch, max_timesteps
    return cfg


class SamplerPlan:
    def __init__(self, model, device, precision: Optional[str] = None, max_batch: Optional[int] = None,
                 max_timesteps: int = 256):
        self.lib = _capi.load()
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("SamplerPlan needs a CUDA device (sm_100a); there is no CPU path")
        if self.lib.mdt_device_count() == 0:
            raise RuntimeError("no sm_100 CUDA device visible to libmdt_b200.so; there is no CPU path")
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.precision = precision or default_precision()
        self.max_batch = max_batch or default_max_batch(device)
        self.max_timesteps = max_timesteps
        self.sigma_data = model.diffusion.diffusion.sigma_data
        self.pred_dim, self.max_length = model.pred_dim, model.max_length
        self.cfg = make_config(model, self.precision, self.max_batch, max_timesteps)
        sd = {k: v for k, v in model.state_dict().items() if not k.startswith("diffusion.")}
        keep, arr = [], (_capi.MdtTensor * len(sd))()
        for i, (k, v) in enumerate(sd.items()):
            t = v.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            arr[i].name, arr[i].data, arr[i].numel = k.encode(), t.data_ptr(), t.numel()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):   # plan creation must not move the process's current device
            _capi.check(self.lib.mdt_plan_create(C.byref(self.cfg), arr, len(sd), self.index, C.byref(handle)))
        self.handle = handle
        self.ctx_features = model.unet.cfg.context_embedding_features or 0
        self.weights_version = None

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None):
            with torch.cuda.device(self.device):
                self.lib.mdt_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self) -> int:
        return int(self.lib.mdt_plan_device_bytes(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mdt_plan_launch_count(self.handle))

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev32(self, t: torch.Tensor) -> torch.Tensor:
        return t.detach().to(self.device, torch.float32).contiguous()

    # ------------------------------------------------------------------
    def iter_scalars(self, num_steps: int, sigma_schedule, sampler) -> np.ndarray:
        sigmas = sigma_schedule(num_steps, "cpu")
        return build_iter_scalars(sigmas, num_steps, sampler, self.sigma_data)

    def sample(self, sequences, *, noise0=None, step_noise=None, num_steps: int, sigma_schedule, sampler,
               clamp: bool, cond_scale: float, seed: Optional[int] = None, sample_offset: int = 0,
               return_tokens: bool = False, pre_encoded: bool = False):
        """`sequences` is the raw conditioning [B, n] or, with ``pre_encoded``, the encoded embedding [B, n, F]."""
        if num_steps > self.max_timesteps:
            raise ValueError(f"timesteps={num_steps} exceeds this plan's max_timesteps={self.max_timesteps}")
        if self.ctx_features == 0:     # unconditional UNet: `sequences` only carries the batch size (generative.py:862-868)
            sequences, pre_encoded, cond_scale = torch.zeros((sequences.shape[0], 1)), False, 1.0
        if pre_encoded:
            if sequences.dim() != 3 or sequences.shape[2] != self.ctx_features:
                raise ValueError(f"embedding must have shape [B, n, {self.ctx_features}]")
            b, n_ctx = sequences.shape[:2]
        else:
            b, n_ctx = sequences.shape
        P, L = self.pred_dim, self.max_length
        if b == 0:   # an empty shard: nothing to launch (mdt_plan_sample would reject the null data pointer)
            out = torch.empty((0, P, L), dtype=torch.float32, device=self.device)
            return (out, torch.empty((0, L), dtype=torch.uint8, device=self.device)) if return_tokens else out
        table = np.ascontiguousarray(self.iter_scalars(num_steps, sigma_schedule, sampler))
        assert table.shape[1] == len(ITER_SCALAR_FIELDS) and table.dtype == np.float32
        with torch.cuda.device(self.device):
            cond = self._dev32(sequences)
            n0 = self._dev32(noise0) if noise0 is not None else None
            sn = self._dev32(step_noise) if step_noise is not None else None
            if n0 is not None and tuple(n0.shape) != (b, P, L):
                raise ValueError(f"noise must have shape {(b, P, L)}")
            if sn is not None and tuple(sn.shape) != (num_steps - 1, b, P, L):
                raise ValueError(f"step_noise must have shape {(num_steps - 1, b, P, L)}")
            out = torch.empty((b, P, L), dtype=torch.float32, device=self.device)
            tokens = torch.empty((b, L), dtype=torch.uint8, device=self.device) if return_tokens else None
            _capi.check(self.lib.mdt_plan_set_context_mode(self.handle, int(bool(pre_encoded))))
            if isinstance(sampler, KarrasSampler):
                sig = sigma_schedule(num_steps, "cpu")
                s0 = karras_noise_scales(sig, num_steps, sampler)[0]
                _capi.check(self.lib.mdt_plan_set_sampler_mode(self.handle, 1, float(s0), float(sig[0])))
            else:
                _capi.check(self.lib.mdt_plan_set_sampler_mode(self.handle, 0, 0.0, 0.0))
            _capi.check(self.lib.mdt_plan_sample(
                self.handle, cond.data_ptr(), n_ctx, n0.data_ptr() if n0 is not None else None,
                sn.data_ptr() if sn is not None else None, table.ctypes.data, table.shape[0],
                int(seed or 0), int(sample_offset), b, float(cond_scale), int(bool(clamp)),
                out.data_ptr(), tokens.data_ptr() if tokens is not None else None, self._stream()))
            # inputs must outlive the asynchronous launch: tie them to the stream
            for t in (cond, n0, sn):
                if t is not None:
                    t.record_stream(torch.cuda.current_stream(self.device))
        return (out, tokens) if return_tokens else out

    def inpaint(self, sequences, source, mask, *, num_steps: int, num_resamples: int, sigma_schedule, sampler,
                cond_scale: float, noise=None, seed: Optional[int] = None, sample_offset: int = 0):
        """ADPM2Sampler.inpaint (diffusion.py:526-549) on the device: keep `source` where `mask` is True."""
        if num_steps > self.max_timesteps:
            raise ValueError(f"timesteps={num_steps} exceeds this plan's max_timesteps={self.max_timesteps}")
        b, n_ctx = sequences.shape
        P, L = self.pred_dim, self.max_length
        sigmas = sigma_schedule(num_steps, "cpu").detach().to(torch.float32).contiguous()
        table = np.ascontiguousarray(build_iter_scalars(sigmas, num_steps, sampler, self.sigma_data))
        draws = 1 + (num_steps - 1) * 2 * num_resamples
        if tuple(source.shape) != (b, P, L) or tuple(mask.shape) != (b, P, L):
            raise ValueError(f"inpaint and in_paint_mask must have shape {(b, P, L)}")
        with torch.cuda.device(self.device):
            cond, src = self._dev32(sequences), self._dev32(source)
            msk = mask.detach().to(self.device).to(torch.uint8).contiguous()
            nz = self._dev32(noise) if noise is not None else None
            if nz is not None and tuple(nz.shape) != (draws, b, P, L):
                raise ValueError(f"noise must have shape {(draws, b, P, L)}")
            out = torch.empty((b, P, L), dtype=torch.float32, device=self.device)
            sig = sigmas.numpy()
            _capi.check(self.lib.mdt_plan_inpaint(
                self.handle, cond.data_ptr(), n_ctx, src.data_ptr(), msk.data_ptr(), nz.data_ptr() if nz is not None else None,
                table.ctypes.data, sig.ctypes.data, table.shape[0], int(num_resamples), int(seed or 0), int(sample_offset), b,
                float(cond_scale), out.data_ptr(), self._stream()))
            for t in (cond, src, msk, nz):
                if t is not None:
                    t.record_stream(torch.cuda.current_stream(self.device))
        return out

    def unet_forward(self, x, time: float, sequences, cond_scale: float = 1.0, taps=None):
        """One UNetCFG1d evaluation (modules.py:1228-1255) -- kernel-level parity entry point."""
        b, n_ctx = sequences.shape
        with torch.cuda.device(self.device):
            xd, cond = self._dev32(x), self._dev32(sequences)
            out = torch.empty_like(xd)
            if taps is not None:
                _capi.check(self.lib.mdt_plan_enable_taps(self.handle, 1))
            _capi.check(self.lib.mdt_plan_unet_forward(self.handle, xd.data_ptr(), float(time), cond.data_ptr(),
                                                       n_ctx, b, float(cond_scale), out.data_ptr(), self._stream()))
            torch.cuda.synchronize(self.device)
            if taps is not None:
                for name in list(taps):
                    n = self.lib.mdt_plan_read_tap(self.handle, name.encode(), None, 0)
                    if n < 0:
                        taps[name] = None
                        continue
                    buf = np.empty(n, dtype=np.float32)
                    _capi.check(self.lib.mdt_plan_read_tap(self.handle, name.encode(), buf.ctypes.data, n))
                    taps[name] = buf
                _capi.check(self.lib.mdt_plan_enable_taps(self.handle, 0))
        return out
