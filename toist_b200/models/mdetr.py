"""MDETR / TOIST detector and its set criterion behind the reference's API (reference models/mdetr.py:315-1141).

`MDETR.forward(samples, captions, encode_and_save, memory_cache)` keeps the two-phase protocol of engine.py:63-66 and
the `memory_cache` / output dictionaries of models/transformer.py:155-166 and models/mdetr.py:422-462; parameters keep
the reference's names so state dicts are interchangeable.  Everything numerical runs in hand-written sm_100a kernels
(runtime.py); there is no eager / CPU fallback.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from .. import kernels as K
from ..runtime import (BACKBONE, DECODER, ENCODER, HEADS, TEXT, Call, GraphCache, ShadowBank, Spec, Stage, run_stage)
from ..util import dist
from ..util.misc import NestedTensor, h2d
from .backbone import build_backbone
from .matcher import PackedTargets, build_matcher, indices_from_match, match_layers, pack_targets
from .transformer import build_transformer


class MLP(nn.Module):
    """Parameter container of the box head: Linear(256,256) - ReLU - Linear(256,256) - ReLU - Linear(256,4)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))


class ModelRuntime:
    """Lazily built per-model state: bf16 shadow weights and the static description of each autograd stage."""

    def __init__(self):
        self.text_stream = None
        self.text_stream_enabled = os.environ.get("TOIST_TEXT_STREAM", "1") != "0"
        self.bank = ShadowBank()
        self.stages: Optional[Dict[str, Stage]] = None
        self.graphs: Optional[GraphCache] = None
        self.graphs_text: Optional[GraphCache] = None
        self.seed: Optional[torch.Tensor] = None  # device int64 [1]: dropout seed of the current step
        self.grad_sync = None  # set by util.dist.DistributedDataParallel
        self.dirty = True  # parameters may have moved / been reloaded since the shadow bank was built
        self.tok_cache = OrderedDict()
        self.direct = False  # MDETR.enable_direct_grads: stages assign Parameter.grad themselves
        self.anchor: Optional[torch.Tensor] = None

    def __deepcopy__(self, memo):  # the EMA copy (main.py:322) rebuilds its own
        return ModelRuntime()

    def build(self, m: "MDETR") -> None:
        names = [n for n, _ in m.named_parameters()]
        tr = m.transformer
        tcfg = tr.text_encoder.config
        body = m.backbone[0].body

        def pick(*prefixes, exclude=()):
            return [n for n in names if n.startswith(prefixes) and not any(e in n for e in exclude)]

        heads = ("class_embed.", "bbox_embed.", "contrastive_align_projection_")
        self.stages = {
            "backbone": Stage(m, "backbone", pick("backbone.0.body."), prefix="backbone.0.body.", blocks=body.blocks,
                              first_trainable=2 if m.backbone[0].train_backbone else 5,
                              return_interm=m.backbone[0].return_interm_layers),
            "text": Stage(m, "text", pick("transformer.text_encoder.", "transformer.resizer.", exclude=("pooler.",)),
                          prefix="transformer.text_encoder.", resizer_prefix="transformer.resizer.",
                          num_layers=tcfg.num_hidden_layers, num_heads=tcfg.num_attention_heads,
                          eps=float(tcfg.layer_norm_eps), pad_id=int(tcfg.pad_token_id),
                          resizer_p=float(tr.resizer.dropout_p)),
            "encoder": Stage(m, "encoder", pick("input_proj.", "transformer.encoder."), prefix="transformer.encoder.",
                             input_proj_prefix="input_proj.", num_layers=tr.encoder.num_layers, nhead=tr.nhead),
            "decoder": Stage(m, "decoder", pick("transformer.decoder."), prefix="transformer.decoder.",
                             num_layers=tr.decoder.num_layers, nhead=tr.nhead),
            "heads": Stage(m, "heads", pick(*heads), prefix="", contrastive=m.contrastive_align_loss),
        }

    def refresh(self, m: "MDETR", defer_rest: bool = False) -> None:
        if self.stages is None:
            self.build(m)
        sig = self.bank._sig
        # full scan every 64 steps even when nothing flagged a change (e.g. a buffer edited in place by user code)
        self.steps = getattr(self, "steps", 0) + 1
        dirty = self.dirty or self.steps % 64 == 0
        self.dirty = False
        if dirty:
            for st in self.stages.values():
                st.invalidate()  # requires_grad flags may have changed
        self.bank.ensure(m, "backbone.0.body.", m.backbone[0].body, dirty, defer_rest=defer_rest)
        if self.graphs is not None and sig is not None and sig != self.bank._sig:
            self.graphs.clear()  # shadow buffers moved: captured pointers are stale
            self.graphs_text.clear()

    def call(self, name: str, save: bool, drop_p: float = 0.0, **kw) -> Call:
        # the text stage replays concurrently with the backbone: graphs that may overlap in time must not share a
        # memory pool (a pool is only safe for graphs replayed one after the other)
        graphs = self.graphs_text if (name == "text" and self.text_stream_enabled) else self.graphs
        c = Call(self.stages[name], self.bank.w, save, graphs=graphs, drop_p=drop_p,
                 seed=self.seed, **kw)
        c.grad_sync = self.grad_sync  # util.dist.FlatGradSync when wrapped for data-parallel training, else None
        if self.direct:
            dev = next(iter(self.stages[name].params)).device if self.stages[name].params else None
            if self.anchor is None or self.anchor.device != dev:
                self.anchor = torch.zeros(1, device=dev, requires_grad=True)
            c.anchor = self.anchor
        return c

    def tokenize(self, tokenizer, captions, dev):
        """models/transformer.py:129 tokenises the captions on the host every step and copies the ids to the device.
        TOIST's pronoun captions are a small closed set (14 task verbs x 'something', datasets/tdod.py:23-38), so the
        device-resident result is cached by caption tuple (LRU, 512 batches)."""
        key = (tuple(captions), dev)
        hit = self.tok_cache.get(key)
        if hit is not None:
            self.tok_cache.move_to_end(key)
            return hit
        tokenized = tokenizer.batch_encode_plus(list(captions), padding="longest", return_tensors="pt").to(dev)
        ids = tokenized["input_ids"].contiguous()
        attn = tokenized["attention_mask"].to(torch.int64).contiguous()
        hit = (tokenized, ids, attn, attn.ne(1))
        self.tok_cache[key] = hit
        if len(self.tok_cache) > 512:
            self.tok_cache.popitem(last=False)
        return hit

    def new_step_seed(self, device) -> None:
        """Draws the dropout seed of this step from torch's CPU generator (reproducible under torch.manual_seed)."""
        s = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
        if self.seed is None or self.seed.device != device:
            self.seed = h2d(s, device)
        else:
            self.seed.copy_(s.pin_memory(), non_blocking=True)


class MDETR(nn.Module):
    """Modulated detection model: ResNet trunk + RoBERTa -> 6+6 layer cross-modal transformer -> box / token heads."""

    def __init__(self, backbone, transformer, num_classes, num_queries, aux_loss=False, contrastive_hdim=64,
                 contrastive_align_loss=False, cluster_num=16, args=None):
        super().__init__()
        self.args = args
        self.num_queries = num_queries
        self.transformer = transformer
        hidden_dim = transformer.d_model
        self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.input_proj = nn.Conv2d(backbone.num_channels, hidden_dim, kernel_size=1)
        self.backbone = backbone
        self.aux_loss = aux_loss
        self.contrastive_align_loss = contrastive_align_loss
        if contrastive_align_loss:
            self.contrastive_align_projection_image = nn.Linear(hidden_dim, contrastive_hdim)
            self.contrastive_align_projection_text = nn.Linear(hidden_dim, contrastive_hdim)
        self._rt = ModelRuntime()

    # ------------------------------------------------------------------ helpers
    def _drop_probs(self):
        """Training-mode dropout probabilities per stage (the reference's nn.Dropout modules): transformer layers
        `--dropout` (models/transformer.py:273-283,337-353), RoBERTa its config's hidden / attention dropout, the
        resizer 0.1 (models/transformer.py:73).  All zero in eval()."""
        if not self.training:
            return None
        cfg = self.transformer.text_encoder.config
        text_p = float(cfg.hidden_dropout_prob)
        if abs(float(cfg.attention_probs_dropout_prob) - text_p) > 1e-12:
            raise NotImplementedError("RoBERTa with different hidden / attention dropout probabilities")
        return {"text": text_p, "transformer": float(self.transformer.dropout_p)}

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): parameter storage may move
        self._rt.dirty = True
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._rt.dirty = True
        return super().load_state_dict(*args, **kwargs)

    def enable_cuda_graphs(self, on: bool = True) -> "MDETR":
        """Capture each stage's forward / backward launch sequence into CUDA graphs (one per input-shape signature)
        and replay them on later steps.  For fixed-shape training / benchmarking; tensors returned by a step are
        overwritten by the next step with the same shapes."""
        self._rt.graphs = GraphCache(priority=-1) if on else None  # the critical chain; text / wgrad lanes fill around it
        self._rt.graphs_text = GraphCache() if on else None
        return self

    def enable_direct_grads(self, on: bool = True) -> "MDETR":
        """Let every backward stage assign its parameters' `.grad` itself (views of the stage's flat gradient arena)
        instead of returning ~1000 gradients through autograd: removes several ms of host time per step
        (Function.apply over all parameters + one AccumulateGrad node per parameter).  Semantics kept: `.grad` is
        None -> set; `.grad` present -> accumulated.  NOT compatible with torch.nn.parallel.DistributedDataParallel
        (its reducer listens on AccumulateGrad) nor with per-parameter gradient hooks; toist_b200.util.dist.
        DistributedDataParallel switches this mode on by itself."""
        self._rt.direct = bool(on)
        return self

    def _grad_wanted(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    # ------------------------------------------------------------------ phase A (engine.py:63)
    def encode(self, samples: NestedTensor, captions, want_features: bool = False) -> dict:
        rt = self._rt
        # the trunk's shadow weights are refreshed here, in front of the backbone; the rest (RoBERTa, transformer, heads)
        # on the text branch's stream below, next to the backbone, when that stream is in use
        on_text_stream = rt.text_stream_enabled and isinstance(captions[0], str)
        rt.refresh(self, defer_rest=on_text_stream)
        save = self._grad_wanted()
        dp = self._drop_probs()
        images = samples.tensors
        if images.dtype != torch.float32 or not images.is_cuda:
            raise RuntimeError("toist_b200 expects fp32 CUDA images (sm_100a); there is no CPU path")
        images = images.contiguous()
        if dp is not None:
            rt.new_step_seed(images.device)
        else:
            rt.seed = None
        dev = images.device
        E = self.transformer.d_model
        text_stream = None
        if isinstance(captions[0], str):
            tokenized, ids, attn, text_attention_mask = rt.tokenize(self.transformer.tokenizer, captions, dev)
            # The text branch (12 RoBERTa layers on B x L <= a few hundred rows: ~150 tiny, latency-bound launches) does
            # not depend on the image: it runs on its own stream next to the backbone's large kernels, forward and
            # (autograd replays a node's backward on its forward stream) backward.  TOIST_TEXT_STREAM=0 disables.
            main = torch.cuda.current_stream()
            if rt.text_stream_enabled:
                if rt.text_stream is None or rt.text_stream.device != dev:
                    rt.text_stream = torch.cuda.Stream(device=dev)
                text_stream = rt.text_stream
                text_stream.wait_stream(main)
                with torch.cuda.stream(text_stream):
                    rt.bank.run_rest()
                    text_resized = run_stage(TEXT, rt.call("text", save, dp["text"] if dp else 0.0), ids,
                                             text_attention_mask.view(torch.uint8))[0]
            else:
                text_resized = run_stage(TEXT, rt.call("text", save, dp["text"] if dp else 0.0), ids,
                                         text_attention_mask.view(torch.uint8))[0]
        else:  # already encoded (models/transformer.py:139-141)
            text_attention_mask, text_resized, tokenized = captions
            attn = (~text_attention_mask).to(torch.int64).contiguous()
        feats = run_stage(BACKBONE, rt.call("backbone", save), images)
        c5 = feats[-1]
        B, h, w, _ = c5.shape
        if text_stream is not None:
            torch.cuda.current_stream().wait_stream(text_stream)
            text_resized.record_stream(torch.cuda.current_stream())
            for t in (ids, text_attention_mask):
                t.record_stream(text_stream)
        L = text_resized.shape[0]

        pad_u8 = samples.mask.contiguous().view(torch.uint8)
        small, key = K.key_mask(pad_u8, (h, w), attn)
        npf = self.backbone[1].num_pos_feats
        pos32, pos16 = K.pos_sine(small, npf, float(self.backbone[1].temperature), extra_rows=L)
        S = h * w + L
        enc_out = run_stage(ENCODER, rt.call("encoder", save, dp["transformer"] if dp else 0.0), c5, text_resized,
                            pos16.view(S * B, E), key)
        img_memory = enc_out[0]
        query_embed = self.query_embed.weight.unsqueeze(1).repeat(1, B, 1)
        memory_cache = {
            "text_memory_resized": text_resized,
            "text_memory": img_memory[-L:],
            "img_memory": img_memory,
            "text_pooled_op": None,
            "img_pooled_op": None,
            "mask": key.view(torch.bool),
            "text_attention_mask": text_attention_mask,
            "pos_embed": pos32,
            "query_embed": query_embed,
            "tokenized": tokenized,
        }
        if want_features:
            memory_cache["_b200_feats"] = feats
            memory_cache["_b200_src_proj"] = enc_out[1]
            memory_cache["_b200_small_mask"] = small
        return memory_cache

    # ------------------------------------------------------------------ phase B (engine.py:66)
    def decode(self, memory_cache: dict, want_hs: bool = False) -> dict:
        rt = self._rt
        if rt.stages is None:
            rt.refresh(self)
        save = self._grad_wanted()
        dp = self._drop_probs()
        if dp is not None and rt.seed is None:
            rt.new_step_seed(memory_cache["img_memory"].device)
        cluster = bool(getattr(self.args, "cluster", False))
        mem = memory_cache["img_memory_mod"] if cluster else memory_cache["img_memory"]
        S, B, E = mem.shape
        pos16 = K.cast_bf16(memory_cache["pos_embed"].contiguous()).view(S * B, E)
        key = memory_cache["mask"].contiguous().view(torch.uint8)
        hs = run_stage(DECODER, rt.call("decoder", save, dp["transformer"] if dp else 0.0), mem,
                       memory_cache["query_embed"], pos16, key)[0]
        res = run_stage(HEADS, rt.call("heads", save, B=B), hs, memory_cache["text_memory"])
        logits, boxes = res[0], res[1]
        out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1]}
        stacked = {"pred_logits": logits, "pred_boxes": boxes}
        if self.contrastive_align_loss:
            pq, ptok = res[2], res[3]
            out.update({"proj_queries": pq[-1], "proj_tokens": ptok, "tokenized": memory_cache["tokenized"]})
            stacked.update({"proj_queries": pq, "proj_tokens": ptok})
        if self.aux_loss:
            aux = []
            for l in range(logits.shape[0] - 1):
                a = {"pred_logits": logits[l], "pred_boxes": boxes[l]}
                if self.contrastive_align_loss:
                    a.update({"proj_queries": stacked["proj_queries"][l], "proj_tokens": stacked["proj_tokens"],
                              "tokenized": memory_cache["tokenized"]})
                aux.append(a)
            out["aux_outputs"] = aux
        out["_b200_stacked"] = stacked  # all decoder layers in one tensor each: the criterion consumes these
        if want_hs:
            out["_b200_hs"] = hs
        return out

    def forward(self, samples: NestedTensor, captions, encode_and_save=True, memory_cache=None):
        if not isinstance(samples, NestedTensor):
            samples = NestedTensor.from_tensor_list(samples)
        if encode_and_save:
            assert memory_cache is None
            return self.encode(samples, captions)
        assert memory_cache is not None
        return self.decode(memory_cache)


# ====================================================================================================== criterion
def _criterion_fwd(c: Call, logits, boxes, pq, ptok, tgt_boxes, tgt_count, posmap, tok_pos, num_boxes, forced=None):
    """Matching + every detection loss term of every decoder layer in a handful of launches; returns out [5, L] (see
    toist_criterion_reduce), the assignments and the error flag.  Unit gradients w.r.t. logits, boxes and both
    contrastive projections are produced by the same launches (loss_contrastive_align is differentiable in the
    reference, models/mdetr.py:601-666; only loss_cardinality :783 is evaluated without gradient)."""
    pt = PackedTargets(tgt_boxes, tgt_count, posmap, (), tgt_boxes.shape[1])
    if forced is None:
        match_q, flags, _ = match_layers(logits, boxes, pt, c.w_class, c.w_bbox, c.w_giou)
    else:  # SetCriterion.force_match: differentiate a given assignment (parity tests against reference goldens)
        match_q, flags = forced, torch.zeros(1, dtype=torch.int32, device=logits.device)
    row_loss, dlogits = K.token_ce(logits, match_q, tgt_count, posmap, num_boxes, c.eos_coef, c.save)
    pl1, pgi, d1, d2 = K.box_loss(boxes, match_q, tgt_count, tgt_boxes, num_boxes, c.save)
    card = K.cardinality(logits)
    img_loss = dpq = dpt = None
    if pq is not None:
        img_loss, dpq, dpt = K.contrastive_align(pq, ptok, match_q, tgt_count, tok_pos, num_boxes, c.temperature,
                                                 c.save)
    out = K.criterion_reduce(row_loss, pl1, pgi, card, img_loss, tgt_count, num_boxes, flags)
    return (out, match_q, flags), ((dlogits, d1, d2, dpq, dpt) if c.save else None)


def _criterion_bwd(c: Call, saved, needs, gout, *unused):
    """gout [5, L]: the weight each (term, layer) cell carries in the caller's weighted sum (engine.py:72)."""
    dlogits, d1, d2, dpq, dpt = saved
    gl = K.scale_layers(dlogits, gout[0]) if needs[0] else None
    gb = K.scale_layers2(d1, gout[1], d2, gout[2]) if needs[1] else None
    gq = gt = None
    if dpq is not None:
        g4 = gout[4].contiguous()
        gq = K.scale_layers(dpq, g4) if needs[2] else None
        gt = K.scale_layers(dpt, g4, reduce=True) if needs[3] else None  # proj_tokens is shared by all decoder layers
    return (gl, gb, gq, gt) + (None,) * 6, {}


CRITERION = Spec("criterion", 10, _criterion_fwd, _criterion_bwd, nondiff=(1, 2))


class _LossTerms(torch.autograd.Function):
    """out [5, L] -> its 5 * L scalars as separate 0-dim tensors (the reference's loss dict holds one tensor per term
    and layer, engine.py:70-76 sums them with their weights).  Indexing `out[row, l]` thirty times would hand autograd
    thirty SelectBackward nodes, each materialising a zero [5, L] tensor that is then accumulated pairwise (~90 tiny
    launches in front of the backward pass); here the thirty incoming scalar gradients are gathered by ONE stack."""

    @staticmethod
    def forward(ctx, out):
        ctx.shape = out.shape
        return tuple(out.detach().reshape(-1).unbind(0))

    @staticmethod
    def backward(ctx, *grads):
        ref = next(g for g in grads if g is not None)
        zero = None
        cols = []
        for g in grads:
            if g is None:
                if zero is None:
                    zero = torch.zeros((), dtype=ref.dtype, device=ref.device)
                g = zero
            cols.append(g.reshape(()))
        return torch.stack(cols).view(ctx.shape)


def _mask_loss_fwd(c: Call, pred_masks, tgt_masks, match_q, tgt_count, num_boxes):
    """loss_masks (models/mdetr.py:827-853): matched predictions, bilinear upsample to the padded target size,
    sigmoid focal + dice, fused in one pass; returns [loss_mask, loss_dice]."""
    out, sums = K.mask_loss_fwd(pred_masks, tgt_masks, match_q, tgt_count, num_boxes)
    return (out,), ((pred_masks, tgt_masks, match_q, tgt_count, sums, num_boxes) if c.save else None)


def _mask_loss_bwd(c: Call, saved, needs, gout):
    pred_masks, tgt_masks, match_q, tgt_count, sums, num_boxes = saved
    dpred = K.mask_loss_bwd(pred_masks, tgt_masks, match_q, tgt_count, sums, num_boxes, gout.contiguous())
    return (dpred, None, None, None, None), {}


MASKLOSS = Spec("mask_loss", 5, _mask_loss_fwd, _mask_loss_bwd)


def _softkd_fwd(c: Call, logits_n, logits_s, boxes_n, boxes_s, match_n, match_s, tgt_count):
    """loss_softkd of every decoder layer (models/mdetr.py:543-599); NaN / infeasible matching costs poison the loss
    (the reference's scipy call raises ValueError) without a device synchronisation."""
    flags = torch.zeros(1, dtype=torch.int32, device=logits_s.device)
    loss, ws = K.softkd_fwd(logits_n, logits_s, boxes_n, boxes_s, match_n, match_s, tgt_count, flags)
    bi_n, bi_s, pair, n_fp = ws[:4]
    saved = (logits_s, bi_n, bi_s, pair, tgt_count, n_fp, int(match_s.shape[-1])) if c.save else None
    return (loss, flags), saved


def _softkd_bwd(c: Call, saved, needs, gout, *unused):
    logits_s, bi_n, bi_s, pair, tgt_count, n_fp, tmax = saved
    d = K.softkd_bwd(logits_s, bi_n, bi_s, pair, tgt_count, n_fp, gout.contiguous(), tmax) if needs[1] else None
    return (None, d, None, None, None, None, None), {}


SOFTKD = Spec("softkd", 7, _softkd_fwd, _softkd_bwd, nondiff=(1,))


def pack_target_masks(targets, t_max: int, device) -> torch.Tensor:
    """uint8 [B, t_max, H, W]: every image's target masks padded to the largest height / width of the batch
    (NestedTensor.from_tensor_list in the reference, models/mdetr.py:840)."""
    H = max(int(t["masks"].shape[-2]) for t in targets)
    W = max(int(t["masks"].shape[-1]) for t in targets)
    out = torch.zeros((len(targets), t_max, H, W), dtype=torch.uint8, device=device)
    for i, t in enumerate(targets):
        m = t["masks"]
        if m.shape[0]:
            out[i, : m.shape[0], : m.shape[-2], : m.shape[-1]] = m.to(device=device, dtype=torch.uint8)
    return out


def _token_spans(tokenized, i: int, spans):
    """char span -> token span with the reference's fallbacks (models/mdetr.py:622-643)."""
    res = []
    for (beg, end) in spans:
        beg_pos = tokenized.char_to_token(i, beg)
        end_pos = tokenized.char_to_token(i, end - 1)
        if beg_pos is None:
            try:
                beg_pos = tokenized.char_to_token(beg + 1)
                if beg_pos is None:
                    beg_pos = tokenized.char_to_token(beg + 2)
            except Exception:
                beg_pos = None
        if end_pos is None:
            try:
                end_pos = tokenized.char_to_token(end - 2)
                if end_pos is None:
                    end_pos = tokenized.char_to_token(end - 3)
            except Exception:
                end_pos = None
        if beg_pos is None or end_pos is None:
            continue
        res.append((beg_pos, end_pos))
    return res


def build_token_positive(tokenized, targets, t_max: int, n_tokens: int) -> torch.Tensor:
    """uint8 [B, t_max, n_tokens]: token j belongs to a positive span of target t of image b.  Built once per batch on
    the host (the reference rebuilds it per decoder layer inside the loss, models/mdetr.py:614-645)."""
    B = len(targets)
    tp = torch.zeros((B, t_max, n_tokens), dtype=torch.uint8)
    for i, tgt in enumerate(targets):
        spans_all = tgt["tokens_positive"] if "tokens_positive" in tgt else tgt["tokens"]
        for t in range(min(len(spans_all), t_max)):
            for (b, e) in _token_spans(tokenized, i, spans_all[t]):
                tp[i, t, b: e + 1] = 1
    return tp


class SetCriterion(nn.Module):
    """Hungarian matching + soft-token CE, L1 / GIoU box losses, cardinality error and contrastive alignment for the
    last decoder layer and every auxiliary layer, evaluated in a handful of launches without host synchronisation."""

    _TERMS = ("loss_ce", "loss_bbox", "loss_giou", "cardinality_error", "loss_contrastive_align")

    def __init__(self, args, num_classes, matcher, eos_coef, losses, temperature, contrastive_hdim, task_count=14):
        super().__init__()
        self.args = args
        self.num_classes = num_classes
        self.matcher = matcher
        self.eos_coef = eos_coef
        self.losses = losses
        self.temperature = temperature
        unsupported = [l for l in losses if l not in ("labels", "boxes", "cardinality", "contrastive_align", "masks",
                                                      "nsthl2", "softkd")]
        self._unsupported = unsupported
        self._stage = Stage.empty("criterion")
        self._stage_mask = Stage.empty("mask_loss")
        self._stage_kd = Stage.empty("softkd")
        self._graphs: Optional[GraphCache] = None
        self._forced = None
        self._fused_sum = False

    def force_match(self, indices) -> None:
        """Testing hook: `indices[l][b] = (query_idx, target_idx)` (the matcher's output format, one list per decoder
        layer, first layer first) replaces the Hungarian assignment in the following forward calls; `None` restores
        the matcher.  Lets a gradient-parity test differentiate exactly the assignment the reference chose when bf16
        input noise would flip near-tie costs."""
        self._forced = indices

    def _forced_match(self, L: int, packed, dev) -> torch.Tensor:
        mq = torch.full((L, len(packed.counts), packed.t_max), -1, dtype=torch.int32)
        for l in range(L):
            for b, (qi, ti) in enumerate(self._forced[l]):
                mq[l, b, ti.to(torch.int64)] = qi.to(torch.int32)
        return h2d(mq, dev)

    def enable_fused_loss_sum(self, on: bool = True) -> "SetCriterion":
        """Hand out the loss terms as `LossValue`s (models/lossvalue.py): the caller's `sum(loss_dict[k] * weight_dict[k]
        ...)` and `.backward()` (engine.py:72,88) then cost one small upload instead of ~75 tiny launches; any other use
        of a term behaves like the plain tensor it stands for."""
        self._fused_sum = bool(on)
        return self

    def enable_cuda_graphs(self, on: bool = True) -> "SetCriterion":
        """Replay the criterion's launch sequence as a CUDA graph (fixed shapes; see MDETR.enable_cuda_graphs)."""
        self._graphs = GraphCache() if on else None
        return self

    def __deepcopy__(self, memo):
        import copy

        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, None if k == "_graphs" else copy.deepcopy(v, memo))
        return new

    def _stack(self, outputs: dict) -> Dict[str, torch.Tensor]:
        st = outputs.get("_b200_stacked")
        if st is not None:
            return st
        layers = list(outputs.get("aux_outputs", [])) + [outputs]
        res = {k: torch.stack([o[k] for o in layers]) for k in ("pred_logits", "pred_boxes")}
        if "proj_queries" in outputs:
            res["proj_queries"] = torch.stack([o["proj_queries"] for o in layers])
            res["proj_tokens"] = outputs["proj_tokens"]
        return res

    def num_boxes_tensor(self, targets, device) -> torch.Tensor:
        """Global mean number of target boxes, clamped to >= 1, kept on the device (models/mdetr.py:997-1001 does an
        all-reduce followed by a blocking .item())."""
        n = sum(len(t["labels"]) for t in targets)
        nb = h2d(torch.as_tensor([n], dtype=torch.float), device)
        if dist.is_dist_avail_and_initialized():
            torch.distributed.all_reduce(nb)
        return torch.clamp(nb / dist.get_world_size(), min=1)

    def forward(self, memory_cache, outputs, targets, positive_map, example_rel=None):
        if self._unsupported:
            raise NotImplementedError(f"losses {self._unsupported} are not built yet")
        if isinstance(outputs, list):
            return self._forward_distillation(memory_cache, outputs, targets, positive_map)
        losses, _ = self._forward_single(outputs, targets, positive_map)
        return losses

    def _forward_single(self, outputs, targets, positive_map, prefix: str = ""):
        """models/mdetr.py:990-1021 (and, with `prefix`, one half of the distillation branch :893-951)."""
        st = self._stack(outputs)
        logits, boxes = st["pred_logits"], st["pred_boxes"]
        if logits.dtype != torch.float32 or not logits.is_cuda:
            raise RuntimeError("toist_b200 criterion expects fp32 CUDA predictions; there is no CPU path")
        L = logits.shape[0]
        dev = logits.device
        use_aux = "aux_outputs" in outputs
        packed = pack_targets(targets, positive_map, dev)
        nb = self.num_boxes_tensor(targets, dev)
        pq = ptok = tok_pos = None
        if "contrastive_align" in self.losses:
            pq, ptok = st["proj_queries"], st["proj_tokens"]
            tok_pos = h2d(build_token_positive(outputs["tokenized"], targets, packed.t_max, ptok.shape[1]), dev)
        save = torch.is_grad_enabled() and (logits.requires_grad or boxes.requires_grad
                                            or (pq is not None and (pq.requires_grad or ptok.requires_grad)))
        # `tag`: the distillation branch evaluates this stage twice per step (noun / sth); each pass needs its own
        # captured graph, a shared one would overwrite the first pass's outputs and saved tensors on its second replay
        call = Call(self._stage, {}, save, graphs=self._graphs, w_class=float(self.matcher.cost_class),
                    w_bbox=float(self.matcher.cost_bbox), w_giou=float(self.matcher.cost_giou),
                    eos_coef=float(self.eos_coef), temperature=float(self.temperature), tag=prefix)
        forced = self._forced_match(L, packed, dev) if self._forced is not None else None
        out, match_q, flags = run_stage(CRITERION, call, logits, boxes, pq, ptok, packed.boxes, packed.count,
                                        packed.posmap, tok_pos, nb, forced)
        self.last_match = (match_q, packed.counts, flags)
        terms = [("loss_ce", 0), ("loss_bbox", 1), ("loss_giou", 2), ("cardinality_error", 3)]
        if pq is not None:
            terms.append(("loss_contrastive_align", 4))
        mask_out = None
        if "masks" in self.losses:  # last decoder layer only (models/mdetr.py:1013-1015)
            assert "pred_masks" in outputs
            pm_ = outputs["pred_masks"]
            tgt_masks = pack_target_masks(targets, packed.t_max, dev)
            msave = torch.is_grad_enabled() and pm_.requires_grad
            mcall = Call(self._stage_mask, {}, msave, graphs=self._graphs, tag=prefix)
            mask_out = run_stage(MASKLOSS, mcall, pm_, tgt_masks, match_q[L - 1].contiguous(), packed.count, nb)[0]
        losses = {}
        if self._fused_sum and out.requires_grad:
            from .lossvalue import loss_cells

            # rows: ce, bbox, giou, cardinality (no gradient, mdetr.py:783), contrastive align
            cells = loss_cells(out, (True, True, True, False, True))[1]
        else:
            cells = _LossTerms.apply(out) if out.requires_grad else tuple(out.reshape(-1).unbind(0))
        for name, row in terms:
            if name == "loss_contrastive_align" and mask_out is not None:
                losses[prefix + "loss_mask"], losses[prefix + "loss_dice"] = mask_out[0], mask_out[1]
                mask_out = None
            v = cells[row * L + L - 1]
            losses[prefix + name] = v.detach() if row == 3 else v  # cardinality_error: no gradient (mdetr.py:783)
        if mask_out is not None:
            losses[prefix + "loss_mask"], losses[prefix + "loss_dice"] = mask_out[0], mask_out[1]
        if use_aux:
            for i in range(L - 1):
                for name, row in terms:
                    v = cells[row * L + i]
                    losses[f"{prefix}{name}_{i}"] = v.detach() if row == 3 else v
        if "nsthl2" in self.losses and not prefix:  # single model: the term is a constant zero (models/mdetr.py:669-670)
            zero = torch.zeros((), device=dev)
            losses["loss_nsthl2"] = zero
            if use_aux:
                for i in range(L - 1):
                    losses[f"loss_nsthl2_{i}"] = zero
        return losses, (st, match_q, packed, flags)

    def _forward_distillation(self, memory_cache, outputs, targets, positive_map):
        """models/mdetr.py:887-989: teacher (noun) and student (pronoun) losses with `noun_` / `sth_` prefixes, then the
        soft-KD term of every decoder layer."""
        outputs_noun, outputs_sth = outputs
        targets_noun, targets_sth = targets
        pm_noun, pm_sth = positive_map
        losses, (st_n, mq_n, pk_n, fl_n) = self._forward_single(outputs_noun, targets_noun, pm_noun, "noun_")
        l_sth, (st_s, mq_s, pk_s, fl_s) = self._forward_single(outputs_sth, targets_sth, pm_sth, "sth_")
        losses.update(l_sth)
        self.last_match_pair = ((mq_n, pk_n.counts, fl_n), (mq_s, pk_s.counts, fl_s))
        if getattr(self.args, "nsthl2_loss", False):  # models/mdetr.py:974-977: last layer only, no prefix
            losses["loss_nsthl2"] = self._loss_nsthl2(memory_cache, outputs, targets, pk_s.counts)
        if getattr(self.args, "softkd_loss", False):
            if pk_n.counts != pk_s.counts:
                raise ValueError("soft-KD pairs the matched queries of teacher and student by target index: both "
                                 "target lists must have the same length per image (models/mdetr.py:567-571,590-594)")
            ls = st_s["pred_logits"]
            save = torch.is_grad_enabled() and ls.requires_grad
            call = Call(self._stage_kd, {}, save, graphs=self._graphs)
            kd, kd_flags = run_stage(SOFTKD, call, st_n["pred_logits"].detach(), ls, st_n["pred_boxes"].detach(),
                                     st_s["pred_boxes"].detach(), mq_n, mq_s, pk_s.count)
            self.last_softkd_flags = kd_flags
            L = ls.shape[0]
            losses["loss_softkd"] = kd[L - 1]
            if "aux_outputs" in outputs_sth:
                for i in range(L - 1):
                    losses[f"loss_softkd_{i}"] = kd[i]
        return losses

    def _loss_nsthl2(self, memory_cache, outputs, targets, counts_sth) -> torch.Tensor:
        """models/mdetr.py:668-781.  Per model (teacher: noun caption, student: pronoun caption) and image, the mean over
        the boxes of the mean `text_memory` row over the box's `noun_tokens_positive` tokens; the loss is the MSE between
        the student's and the (detached) teacher's vector, averaged over the images whose student assignment is not
        empty (= images with at least one target).  An image without boxes keeps the zero vector (:695-697), a box whose
        spans select no token makes the mean NaN, as in the reference."""
        from .cluster import _MseRows, _selection, _TokenWeightedSum  # (cluster.py imports this module)

        feats = []
        for mc, out, tg in zip(memory_cache, outputs, targets):
            text = mc["text_memory"]  # [T, B, D]
            T, bs, _ = text.shape
            w = np.zeros((bs, T), dtype=np.float32)
            for i, t in enumerate(tg):
                per_box = _selection(out["tokenized"], i, t["noun_tokens_positive"], T)
                for pos in per_box:
                    n = pos.sum()
                    w[i] += (pos / n if n > 0 else np.full(T, np.nan, np.float32)) / len(per_box)
            feats.append(_TokenWeightedSum.apply(text, h2d(torch.from_numpy(w), text.device)))
        noun, sth = feats
        dev = sth.device
        use = [1 if c > 0 else 0 for c in counts_sth]  # len(indices_sth[i][0]) = min(Q, T_i) > 0
        if sum(use) == 0:
            return torch.zeros((), device=dev)
        return _MseRows.apply(sth, noun.detach(), h2d(torch.tensor(use, dtype=torch.uint8), dev))

    def last_indices(self) -> List[List[tuple]]:
        """Assignments of the most recent forward as the reference's index pairs, one list per decoder layer
        (synchronises; for tests and debugging)."""
        match_q, counts, flags = self.last_match
        if int(flags.item()) != 0:
            raise ValueError("matrix contains invalid numeric entries")
        mq = match_q.cpu()
        return [indices_from_match(mq[l], counts) for l in range(mq.shape[0])]


def build(args):
    num_classes = 255
    device = torch.device(args.device)
    assert not args.masks or args.mask_model != "none"
    backbone = build_backbone(args)
    transformer = build_transformer(args)
    model = MDETR(backbone, transformer, num_classes=num_classes, num_queries=args.num_queries, aux_loss=args.aux_loss,
                  contrastive_hdim=args.contrastive_loss_hdim, contrastive_align_loss=args.contrastive_align_loss,
                  cluster_num=args.cluster_num, args=args)
    if args.mask_model != "none":
        from .segmentation import DETRsegm

        model = DETRsegm(model, mask_head=args.mask_model, freeze_detr=(args.frozen_weights is not None))
    matcher = build_matcher(args)
    weight_dict = {"loss_ce": args.ce_loss_coef, "loss_bbox": args.bbox_loss_coef}
    if args.contrastive_align_loss:
        weight_dict["loss_contrastive_align"] = args.contrastive_align_loss_coef
    if args.nsthl2_loss:
        weight_dict["loss_nsthl2"] = args.nsthl2_coef
    if args.softkd_loss:
        weight_dict["loss_softkd"] = args.softkd_coef
    if args.cluster and args.distillation:
        weight_dict["loss_cluster_choice"] = args.cluster_choice_loss
        weight_dict["loss_cluster_feature"] = args.cluster_feature_loss
    weight_dict["loss_giou"] = args.giou_loss_coef
    if args.masks:
        weight_dict["loss_mask"] = args.mask_loss_coef
        weight_dict["loss_dice"] = args.dice_loss_coef

    def with_aux(d):
        if args.aux_loss:
            extra = {}
            for i in range(args.dec_layers - 1):
                extra.update({k + f"_{i}": v for k, v in d.items()})
            d.update(extra)
        return d

    if args.distillation:
        cross = ("loss_nsthl2", "loss_softkd", "loss_cluster_choice", "loss_cluster_feature")
        prefixed = {}
        for k, v in weight_dict.items():
            if k in cross:
                prefixed[k] = v
            else:
                prefixed["noun_" + k] = v
                prefixed["sth_" + k] = v
        weight_dict = with_aux(prefixed)
    else:
        weight_dict = with_aux(weight_dict)

    losses = ["labels", "boxes", "cardinality"]
    if args.masks:
        losses += ["masks"]
    if args.contrastive_align_loss:
        losses += ["contrastive_align"]
    if args.nsthl2_loss:
        losses += ["nsthl2"]
    if args.softkd_loss:
        losses += ["softkd"]
    criterion = SetCriterion(args, num_classes, matcher=matcher, eos_coef=args.eos_coef, losses=losses,
                             temperature=args.temperature_NCE, contrastive_hdim=args.contrastive_loss_hdim,
                             task_count=14)
    criterion.to(device)
    cluster_criterion = None
    if args.cluster:
        from .cluster import ClusterCriterion

        cluster_criterion = ClusterCriterion(feature_dim=args.hidden_dim, memory_size=args.cluster_memory_size,
                                             cluster_num=args.cluster_num, task_count=14, args=args)
    return model, criterion, cluster_criterion, weight_dict
