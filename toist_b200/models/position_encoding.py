"""Sine position embedding (reference models/position_encoding.py:13-49, normalize=True) on the sm_100a kernel."""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import kernels as K


class PositionEmbeddingSine(nn.Module):
    """Normalised 2-D sine/cosine embedding, `num_pos_feats` channels per axis (y first, then x)."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if not normalize:
            raise NotImplementedError("toist_b200 implements the normalize=True embedding the reference builds "
                                      "(models/position_encoding.py:92)")
        if scale is not None and abs(scale - 2 * math.pi) > 1e-9:
            raise NotImplementedError("only scale = 2*pi (the reference default) is implemented")
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi

    @torch.no_grad()
    def forward(self, tensor_list):
        """NestedTensor with mask [B,H,W] -> pos [B, 2*num_pos_feats, H, W] fp32 (the reference's layout)."""
        mask = tensor_list.mask
        b, h, w = mask.shape
        p32, _ = K.pos_sine(mask.contiguous().view(torch.uint8), self.num_pos_feats, float(self.temperature))
        return p32.view(h, w, b, -1).permute(2, 3, 0, 1)


def build_position_encoding(args):
    n_steps = args.hidden_dim // 2
    if args.position_embedding in ("v2", "sine"):
        return PositionEmbeddingSine(n_steps, normalize=True)
    raise ValueError(f"position embedding {args.position_embedding!r} is outside the TOIST hot path (sine only)")
