"""Cross-modal transformer: parameter containers with the reference's module / state-dict names
(reference models/transformer.py:22-96,270-349,473-525).  The arithmetic lives in runtime.TEXT / ENCODER /
DECODER (tcgen05 GEMMs, fused epilogues); these classes only own parameters, the tokenizer and hyper-parameters.
"""
from __future__ import annotations

import copy

import torch
from torch import nn

from ..tokenizer import build_tokenizer


def _clones(module: nn.Module, n: int) -> nn.ModuleList:
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


class TransformerEncoderLayer(nn.Module):
    """Post-norm layer: self-attention (q = k = src + pos, v = src), FFN 256 -> 2048 -> 256 with ReLU."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        if normalize_before or activation != "relu":
            raise NotImplementedError("the TOIST hot path is post-norm with ReLU (main.py:145, transformer.py:32)")
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout_p = dropout


class TransformerDecoderLayer(nn.Module):
    """Post-norm layer: self-attention, cross-attention over the image+text memory, FFN (norm1 / norm3 / norm4)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        if normalize_before or activation != "relu":
            raise NotImplementedError("the TOIST hot path is post-norm with ReLU (main.py:145, transformer.py:32)")
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.cross_attn_image = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.norm4 = nn.LayerNorm(d_model)
        self.dropout_p = dropout


class TransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = _clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False):
        super().__init__()
        self.layers = _clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate


class FeatureResizer(nn.Module):
    """Linear 768 -> d_model followed by LayerNorm(eps=1e-12) (and dropout in training)."""

    def __init__(self, input_feat_size, output_feat_size, dropout, do_ln=True):
        super().__init__()
        if not do_ln:
            raise NotImplementedError("FeatureResizer without LayerNorm is not on the TOIST hot path")
        self.do_ln = do_ln
        self.fc = nn.Linear(input_feat_size, output_feat_size, bias=True)
        self.layer_norm = nn.LayerNorm(output_feat_size, eps=1e-12)
        self.dropout_p = dropout


class Transformer(nn.Module):
    def __init__(self, args=None, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6,
                 dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False,
                 return_intermediate_dec=False, pass_pos_and_query=True, text_encoder_type="roberta-base",
                 freeze_text_encoder=False, contrastive_loss=False):
        super().__init__()
        self.args = args
        if not pass_pos_and_query:
            raise NotImplementedError("pass_pos_and_query=False is outside the TOIST hot path (main.py:146-151)")
        if contrastive_loss:
            raise NotImplementedError("the CLS-token contrastive loss is off in every TOIST recipe (main.py:178)")
        if not return_intermediate_dec:
            raise NotImplementedError("the reference always builds the decoder with return_intermediate=True")
        self.pass_pos_and_query = pass_pos_and_query
        enc_layer = TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.encoder = TransformerEncoder(enc_layer, num_encoder_layers, None)
        dec_layer = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.decoder = TransformerDecoder(dec_layer, num_decoder_layers, nn.LayerNorm(d_model),
                                          return_intermediate=return_intermediate_dec)
        self.CLS = None
        self._reset_parameters()

        from transformers import RobertaConfig, RobertaModel

        synthetic = getattr(args, "synthetic_tokenizer", None)
        self.tokenizer = build_tokenizer(text_encoder_type, synthetic)
        if getattr(args, "without_pretrain", False):
            self.text_encoder = RobertaModel(RobertaConfig(type_vocab_size=1, vocab_size=50265))
        else:
            self.text_encoder = RobertaModel.from_pretrained(text_encoder_type)
        if freeze_text_encoder:
            for p in self.text_encoder.parameters():
                p.requires_grad_(False)
        self.expander_dropout = 0.1
        config = self.text_encoder.config
        self.resizer = FeatureResizer(input_feat_size=config.hidden_size, output_feat_size=d_model,
                                      dropout=self.expander_dropout)
        self.d_model = d_model
        self.nhead = nhead
        self.dropout_p = dropout

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


def build_transformer(args):
    return Transformer(
        args=args,
        d_model=args.hidden_dim,
        dropout=args.dropout,
        nhead=args.nheads,
        dim_feedforward=args.dim_feedforward,
        num_encoder_layers=args.enc_layers,
        num_decoder_layers=args.dec_layers,
        normalize_before=args.pre_norm,
        return_intermediate_dec=True,
        pass_pos_and_query=args.pass_pos_and_query,
        text_encoder_type=args.text_encoder_type,
        freeze_text_encoder=args.freeze_text_encoder,
        contrastive_loss=args.contrastive_loss,
    )
