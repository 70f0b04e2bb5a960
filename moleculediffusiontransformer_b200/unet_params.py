"""Parameter container for the 1-D classifier-free-guidance UNet denoiser.

This module holds *parameters only*.  It reproduces the reference's ``state_dict`` key
layout and its parameter construction order (so that ``torch.manual_seed(s)`` followed by
construction yields bit-identical random-init weights), but it contains no forward pass:
the forward pass is the CUDA plan behind ``csrc/`` (see ``plan.py``).

Key layout / construction order follow the reference modules:
  UNet1d.__init__            modules.py:935-1098   (to_mapping, to_time, to_in, downsamples,
                                                     bottleneck, upsamples, to_out)
  UNetCFG1d.__init__         modules.py:1215-1226  (+ fixed_embedding)
  DownsampleBlock1d.__init__ modules.py:575-678    (pre_transformer_block, downsample, blocks, transformer)
  UpsampleBlock1d.__init__   modules.py:725-818    (pre_transformer_block, blocks, transformer, upsample)
  BottleneckBlock1d.__init__ modules.py:866-915
  ResnetBlock1d.__init__     modules.py:146-191    (block1, to_scale_shift, block2, to_out)
  Transformer1d.__init__     modules.py:470-517    (to_in = GN32 + 1x1 conv, blocks, to_out = 1x1 conv)
  TransformerBlock/Attention modules.py:419-447, 368-399
"""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Optional, Sequence

import torch
import torch.nn as nn


@dataclass
class UNetConfig:
    """Static description of one UNet; mirrors the kwargs of ``XUNet1d`` (modules.py:1316)."""

    in_channels: int
    channels: int
    multipliers: Sequence[int]
    factors: Sequence[int]
    num_blocks: Sequence[int]
    attentions: Sequence[int]
    patch_size: int = 1
    resnet_groups: int = 8
    kernel_multiplier_downsample: int = 2
    use_skip_scale: bool = True
    out_channels: Optional[int] = None
    context_features_multiplier: int = 4
    context_embedding_features: Optional[int] = 128     # None: unconditional UNet1d (type='base'), no cross-attention
    context_embedding_max_length: Optional[int] = 12
    pre_transformer: int = 0
    attention_heads: int = 8
    attention_features: int = 64
    attention_multiplier: int = 2
    attention_use_rel_pos: bool = False

    def __post_init__(self):
        self.multipliers = list(self.multipliers)
        self.factors = list(self.factors)
        self.num_blocks = list(self.num_blocks)
        self.attentions = list(self.attentions)
        if self.out_channels is None:
            self.out_channels = self.in_channels
        n = len(self.multipliers) - 1
        # same consistency assertion as modules.py:988-992
        assert len(self.factors) == n and len(self.attentions) >= n and len(self.num_blocks) == n
        if self.attention_use_rel_pos:
            raise NotImplementedError("relative position bias is disabled on the QM path (generative.py:773)")

    @property
    def num_levels(self) -> int:
        return len(self.multipliers) - 1

    @property
    def mapping_features(self) -> int:
        return self.channels * self.context_features_multiplier

    def to_dict(self):
        return asdict(self)


class _Slot(nn.Identity):
    """Parameter-free placeholder keeping nn.Sequential indices equal to the reference's."""


def _seq(*mods) -> nn.Sequential:
    return nn.Sequential(*mods)


class _Holder(nn.Module):
    """A bare namespace module; children are attached with setattr in reference order."""

    def forward(self, *a, **k):  # pragma: no cover - never a compute path
        raise RuntimeError(
            "moleculediffusiontransformer_b200 parameter containers have no eager forward; "
            "use model.sample(...) which runs the sm_100a CUDA plan"
        )


class _LearnedFourier(_Holder):
    def __init__(self, dim: int):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))  # modules.py:552


def _conv_block(cin: int, cout: int, groups: int) -> _Holder:
    h = _Holder()
    h.groupnorm = nn.GroupNorm(num_groups=groups, num_channels=cin)
    h.project = nn.Conv1d(cin, cout, kernel_size=3, padding=1)
    return h


def _resnet(cin: int, cout: int, groups: int, mapping: Optional[int]) -> _Holder:
    h = _Holder()
    h.block1 = _conv_block(cin, cout, groups)
    if mapping is not None:
        inner = _Holder()
        inner.to_scale_shift = _seq(_Slot(), nn.Linear(mapping, cout * 2))
        h.to_scale_shift = inner
    h.block2 = _conv_block(cout, cout, groups)
    h.to_out = nn.Conv1d(cin, cout, kernel_size=1) if cin != cout else _Slot()
    return h


def _attention(features: int, heads: int, head_features: int, ctx: Optional[int]) -> _Holder:
    mid = heads * head_features
    h = _Holder()
    h.norm = nn.LayerNorm(features)
    h.norm_context = nn.LayerNorm(ctx if ctx else features)
    h.to_q = nn.Linear(features, mid, bias=False)
    h.to_kv = nn.Linear(ctx if ctx else features, mid * 2, bias=False)
    core = _Holder()
    core.to_out = nn.Linear(mid, features)
    h.attention = core
    return h


def _transformer(layers: int, channels: int, cfg: UNetConfig, ctx: Optional[int]) -> _Holder:
    h = _Holder()
    h.to_in = _seq(
        nn.GroupNorm(num_groups=32, num_channels=channels, eps=1e-6, affine=True),
        nn.Conv1d(channels, channels, kernel_size=1),
        _Slot(),
    )
    blocks = []
    for _ in range(layers):
        b = _Holder()
        b.attention = _attention(channels, cfg.attention_heads, cfg.attention_features, None)
        if ctx:
            b.cross_attention = _attention(channels, cfg.attention_heads, cfg.attention_features, ctx)
        mid = channels * cfg.attention_multiplier
        b.feed_forward = _seq(nn.Linear(channels, mid), _Slot(), nn.Linear(mid, channels))
        blocks.append(b)
    h.blocks = nn.ModuleList(blocks)
    h.to_out = _seq(_Slot(), nn.Conv1d(channels, channels, kernel_size=1))
    return h


class UNetCFG1dParams(_Holder):
    """Parameters of ``UNetCFG1d`` (modules.py:1211-1255) under the reference's names."""

    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        c, m = cfg.channels, cfg.mapping_features
        mult, ctx = cfg.multipliers, cfg.context_embedding_features
        nlev = cfg.num_levels

        self.to_mapping = _seq(nn.Linear(m, m), _Slot(), nn.Linear(m, m), _Slot())
        self.to_time = _seq(_seq(_LearnedFourier(c), nn.Linear(c + 1, m)), _Slot())

        assert (c * mult[0]) % cfg.patch_size == 0
        to_in = _Holder()
        to_in.block = _resnet(cfg.in_channels, c * mult[0] // cfg.patch_size, 1, m)
        self.to_in = to_in

        downs = []
        for i in range(nlev):
            cin, cout, f = c * mult[i], c * mult[i + 1], cfg.factors[i]
            d = _Holder()
            if cfg.pre_transformer > 0:
                d.pre_transformer_block = _transformer(cfg.pre_transformer, cout, cfg, None)
            k = f * cfg.kernel_multiplier_downsample + 1
            d.downsample = nn.Conv1d(cin, cout, kernel_size=k, stride=f,
                                     padding=f * (cfg.kernel_multiplier_downsample // 2))
            d.blocks = nn.ModuleList([_resnet(cout, cout, cfg.resnet_groups, m) for _ in range(cfg.num_blocks[i])])
            if cfg.attentions[i] > 0:
                d.transformer = _transformer(cfg.attentions[i], cout, cfg, ctx)
            downs.append(d)
        self.downsamples = nn.ModuleList(downs)

        bott = _Holder()
        cb = c * mult[-1]
        bott.pre_block = _resnet(cb, cb, cfg.resnet_groups, m)
        if cfg.attentions[-1] > 0:
            bott.transformer = _transformer(cfg.attentions[-1], cb, cfg, ctx)
        bott.post_block = _resnet(cb, cb, cfg.resnet_groups, m)
        self.bottleneck = bott

        ups = []
        for i in reversed(range(nlev)):
            cin, cout, f = c * mult[i + 1], c * mult[i], cfg.factors[i]
            u = _Holder()
            if cfg.pre_transformer > 0:
                u.pre_transformer_block = _transformer(cfg.pre_transformer, cin, cfg, None)
            nres = cfg.num_blocks[i] + (1 if cfg.attentions[i] else 0)
            u.blocks = nn.ModuleList([_resnet(cin + cin, cin, cfg.resnet_groups, m) for _ in range(nres)])
            if cfg.attentions[i] > 0:
                u.transformer = _transformer(cfg.attentions[i], cin, cfg, ctx)
            assert f % 2 == 0, "odd upsampling factors are not used by the QM models"
            u.upsample = nn.ConvTranspose1d(cin, cout, kernel_size=f * 2, stride=f, padding=f // 2)
            ups.append(u)
        self.upsamples = nn.ModuleList(ups)

        to_out = _Holder()
        to_out.block = _resnet(c * mult[0] // cfg.patch_size, cfg.out_channels, 1, m)
        self.to_out = to_out

        if ctx:   # UNetCFG1d only (modules.py:1215-1226): the learned null embedding of classifier-free guidance
            fe = _Holder()
            fe.max_length = cfg.context_embedding_max_length
            fe.embedding = nn.Embedding(cfg.context_embedding_max_length, ctx)
            self.fixed_embedding = fe


class UNet1dParams(UNetCFG1dParams):
    """Parameters of the plain ``UNet1d`` (modules.py:934-1098; ``XUNet1d(type='base')``): same construction order, no
    conditioning embedding, hence no cross-attention layers and no ``fixed_embedding``."""

    def __init__(self, cfg: UNetConfig):
        assert not cfg.context_embedding_features, "type='base' takes no context_embedding_features (modules.py:1316-1326)"
        super().__init__(cfg)


def XUNet1d(type: str = "cfg", **kwargs) -> UNetCFG1dParams:
    """Factory with the reference's signature (modules.py:1316-1326): ``type='cfg'`` (UNetCFG1d) or ``'base'`` (UNet1d)."""
    if type == "cfg":
        return UNetCFG1dParams(UNetConfig(**kwargs))
    if type == "base":
        kwargs.setdefault("context_embedding_features", None)
        kwargs.setdefault("context_embedding_max_length", None)
        return UNet1dParams(UNetConfig(**kwargs))
    raise NotImplementedError(f"unet type {type!r} is outside the accelerated path ('cfg' and 'base' only)")
