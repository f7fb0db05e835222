"""Synthetic batches of SURVEY.md §8(d) shared by bench.py, the tests and the golden generator (CPU tensors)."""
from __future__ import annotations

import argparse
from typing import List, Tuple

import torch


def make_args(backbone: str = "resnet50", **over) -> argparse.Namespace:
    """The reference's main.py defaults (main.py:32-274) for the flags the hot path reads, without importing it."""
    d = dict(
        backbone=backbone, dilation=False, position_embedding="sine", masks=False, mask_model="none",
        frozen_weights=None, lr_backbone=1e-5, hidden_dim=256, nheads=8, enc_layers=6, dec_layers=6,
        dim_feedforward=2048, dropout=0.0, num_queries=100, pre_norm=False, pass_pos_and_query=True,
        text_encoder_type="roberta-base", freeze_text_encoder=False, contrastive_loss=False,
        contrastive_align_loss=True, contrastive_loss_hdim=64, temperature_NCE=0.07, set_loss="hungarian",
        set_cost_class=1.0, set_cost_bbox=5.0, set_cost_giou=2.0, ce_loss_coef=1.0, bbox_loss_coef=5.0,
        giou_loss_coef=2.0, mask_loss_coef=1.0, dice_loss_coef=1.0, contrastive_align_loss_coef=1.0, eos_coef=0.1,
        aux_loss=True, nsthl2_loss=False, nsthl2_coef=1.0, softkd_loss=False, softkd_coef=1.0, cluster=False,
        cluster_num=3, cluster_memory_size=1024, cluster_feature_loss=1e4, cluster_choice_loss=0.0,
        distillation=False, train_batch_size=2, fifo_memory=False, without_pretrain=True, device="cuda", synthetic_tokenizer=True,
    )
    d.update(over)
    return argparse.Namespace(**d)


def make_batch(batch: int, size: int, n_tokens: int, seed: int = 1234, pad: bool = False, masks: bool = False):
    """images [B,3,size,size] fp32, pad mask [B,size,size] bool, captions of exactly n_tokens - 2 characters ending
    in 'something', targets (T_i = 1 + i % 4 boxes) and the normalised positive map [sum T_i, 256]."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, size, size, generator=g)
    mask = torch.zeros(batch, size, size, dtype=torch.bool)
    if pad and batch > 1:  # ragged batch: the last image is smaller, padded on the right / bottom
        h, w = size - size // 4, size - size // 8
        images[-1, :, h:, :] = 0
        images[-1, :, :, w:] = 0
        mask[-1, h:, :] = True
        mask[-1, :, w:] = True
    n_chars = n_tokens - 2
    verbs = ["sit", "open", "pour", "dig", "step", "lift", "cut", "hit"]
    captions = []
    for i in range(batch):
        tail = " something"
        head = (verbs[i % len(verbs)] * 8)[: max(n_chars - len(tail), 0)]
        captions.append((head + tail)[-n_chars:] if n_chars >= 1 else "")
    targets = []
    for i in range(batch):
        t = 1 + (i % 4)
        cxcy = torch.rand(t, 2, generator=g) * 0.5 + 0.25
        wh = torch.rand(t, 2, generator=g) * 0.3 + 0.05
        cap = captions[i]
        s = cap.find("something")
        targets.append({
            "boxes": torch.cat([cxcy, wh], -1),
            "labels": torch.ones(t, dtype=torch.long),
            "tokens_positive": [[[0, len(cap)]] for _ in range(t)],
            "noun_tokens_positive": [[[max(s, 0), len(cap)]] for _ in range(t)],
        })
    if masks:  # drawn after everything else so that the detection batches of a given seed do not change
        for i, t in enumerate(targets):
            n = len(t["boxes"])
            h, w = (size - size // 4, size - size // 8) if (pad and batch > 1 and i == batch - 1) else (size, size)
            t["masks"] = torch.rand(n, h, w, generator=g) < 0.5
    total = sum(len(t["boxes"]) for t in targets)
    pm = torch.zeros(total, 256)
    pm[:, 1: n_tokens - 1] = 1.0
    pm = pm / (pm.sum(-1, keepdim=True) + 1e-6)
    return images, mask, captions, targets, pm


def targets_to(targets: List[dict], device) -> List[dict]:
    from .util.misc import h2d  # pinned staging: a pageable copy would drain the stream once per tensor

    return [{k: (h2d(v, device) if isinstance(v, torch.Tensor) else v) for k, v in t.items()} for t in targets]
