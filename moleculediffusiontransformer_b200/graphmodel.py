"""Sibling task wrappers of the reference's ``graphmodel.py``: ``AnalogDiffusionSparse`` / ``AnalogDiffusionFull``.

Same conditioning encoder, same ``XDiffusion_x(type='k')`` front-end and the same sampler call as the QM wrappers
(graphmodel.py:355-389, 547-597); only the UNet hyper-parameters differ (graphmodel.py:264-281, 438-455), so they run on the same
CUDA plan.  ``forward`` (training on padded xyz / neighbour channels, graphmodel.py:316-352) stays with the reference through
``set_training_delegate``.
"""
from __future__ import annotations

from .generative import _QMBase


class AnalogDiffusionSparse(_QMBase):
    """graphmodel.py:225-389: cfg UNet with patch_size 8, two resnets and one transformer layer per level."""

    _default_cond_scale = 7.5
    _unet_kwargs = dict(patch_size=8, multipliers=[1, 2, 4], factors=[4, 4], num_blocks=[2, 2], attentions=[1, 1],
                        attention_heads=8, attention_features=64, attention_multiplier=2, attention_use_rel_pos=False)

    def __init__(self, max_length=1024, channels=128, pred_dim=1, context_embedding_max_length=32, unet_type="cfg",
                 pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=1024, embed_dim_position=64,
                 predict_neighbors=False):
        super().__init__(max_length, channels, pred_dim, None, context_embedding_max_length, unet_type, pos_emb_fourier,
                         pos_emb_fourier_add, text_embed_dim, embed_dim_position)
        self.predict_neighbors = predict_neighbors


class AnalogDiffusionFull(_QMBase):
    """graphmodel.py:392-597: cfg UNet with patch_size 4, three resnets and one transformer layer per level."""

    _default_cond_scale = 7.5
    _unet_kwargs = dict(patch_size=4, multipliers=[1, 2, 4], factors=[4, 4], num_blocks=[3, 3], attentions=[1, 1],
                        attention_heads=8, attention_features=64, attention_multiplier=2, attention_use_rel_pos=False)

    def __init__(self, max_length=1024, channels=128, pred_dim=1, context_embedding_max_length=32, unet_type="cfg",
                 pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=1024, embed_dim_position=64,
                 predict_neighbors=True):
        super().__init__(max_length, channels, pred_dim, None, context_embedding_max_length, unet_type, pos_emb_fourier,
                         pos_emb_fourier_add, text_embed_dim, embed_dim_position)
        self.predict_neighbors = predict_neighbors
