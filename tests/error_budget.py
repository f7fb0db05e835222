"""Per-checkpoint forward error budget: where along the path does the distance to the fp32 oracle come from?

TEST INFRASTRUCTURE (imports oracle/).  For every checkpoint of the forward pass two numbers are printed, both norm-wise
relative errors against the fp32 oracle on the same weights and batch:

  cumulative   the tensor as the real pipeline produces it (all upstream bf16 rounding included)
  stage-local  the same stage fed with the ORACLE's fp32 input (cast to the stage's input format), i.e. what this
               stage alone adds

    python tests/error_budget.py [--backbone resnet50 --batch 2 --size 480 --tokens 8] [--out profiles/x.txt]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from conftest import rel_err  # noqa: E402
from synth import make_args, make_batch  # noqa: E402

BF = torch.bfloat16


def oracle_checkpoints(sd, backbone, batch, tokenizer):
    """fp32 oracle forward on the CPU with every intermediate kept."""
    import torch.nn.functional as F

    from oracle import model as O

    images, mask, captions, _, _ = batch
    cfg = O.Config(backbone=backbone)
    tokd = tokenizer(captions)
    with torch.no_grad():
        mc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
        # encoder / decoder layer by layer (same code as O.encoder / O.decoder, keeping every layer's output)
        src_seq = mc["src_proj"].flatten(2).permute(2, 0, 1)
        x = torch.cat([src_seq, mc["text_memory_resized"]], 0)
        enc = []
        P = "transformer.encoder."
        for i in range(cfg.enc_layers):
            p = f"{P}layers.{i}."
            qk = x + mc["pos_embed"]
            x = O._ln(x + O.mha(qk, qk, x, sd, p + "self_attn.", cfg.nheads, mc["mask"]), sd, p + "norm1.")
            x = O._ln(x + O._ffn(x, sd, p), sd, p + "norm2.")
            enc.append(x)
        out = O.decode(sd, cfg, mc)
    return mc, enc, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--tokens", type=int, default=8)
    ap.add_argument("--out", default="")
    a = ap.parse_args()

    from toist_b200 import blocks as Bk
    from toist_b200 import kernels as K
    from toist_b200 import runtime as R
    from toist_b200.models import build_model
    from toist_b200.util.misc import NestedTensor

    torch.manual_seed(0)
    model, _, _, _ = build_model(make_args(a.backbone))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.cuda().eval()
    data = make_batch(a.batch, a.size, a.tokens, seed=1234, pad=True)
    omc, oenc, oout = oracle_checkpoints(sd, a.backbone, data, model.transformer.tokenizer)
    images, mask, captions = data[0].cuda(), data[1].cuda(), data[2]
    rows = []

    def add(name, cum, loc=None):
        rows.append((name, cum, loc))

    with torch.no_grad():
        mc = model.encode(NestedTensor(images, mask), captions, want_features=True)
        out = model.decode(mc, want_hs=True)
        rt = model._rt
        w = rt.bank.w
        B = images.shape[0]
        E = model.transformer.d_model
        # ---- backbone: layer1..layer4 (NHWC bf16 -> NCHW)
        feats = mc["_b200_feats"]
        if len(feats) == 1:
            add("backbone layer4 (first stage: cumulative = local)", rel_err(feats[0].permute(0, 3, 1, 2).float(),
                                                                              omc["features"][-1]))
        else:
            for i, f in enumerate(feats):
                add(f"backbone layer{i + 1}", rel_err(f.permute(0, 3, 1, 2).float(), omc["features"][i]))
        # ---- text branch
        add("text_memory_resized (RoBERTa + resizer; first stage)", rel_err(mc["text_memory_resized"], omc["text_memory_resized"]))
        # ---- encoder, cumulative per layer: re-run our layers on our own inputs
        c5 = feats[-1]
        _, h, wd, _ = c5.shape
        hw = h * wd
        L = mc["text_memory_resized"].shape[0]
        S = hw + L
        pos16 = K.cast_bf16(mc["pos_embed"].contiguous()).view(S * B, E)
        key = mc["mask"].contiguous().view(torch.uint8)

        def run_encoder(feat_nhwc, text32):
            src = torch.empty((S * B, E), dtype=BF, device="cuda")
            Bk.seq_from_nhwc_fwd(feat_nhwc, w["input_proj.weight"], w["input_proj.bias"], src[: hw * B], B)
            K.cast_bf16(text32.contiguous().view(L * B, E), out=src[hw * B:])
            x = src
            outs = [src.clone()]
            for i in range(model.transformer.encoder.num_layers):
                x = Bk.encoder_layer_fwd(R.WView(w, f"transformer.encoder.layers.{i}."), x, pos16, key,
                                         model.transformer.nhead, B, None)[0]
                outs.append(x)
            return outs

        ours = run_encoder(c5, mc["text_memory_resized"])
        o_c5 = omc["features"][-1].cuda().permute(0, 2, 3, 1).contiguous().to(BF)
        local = run_encoder(o_c5, omc["text_memory_resized"].cuda())
        o_src = torch.cat([omc["src_proj"].flatten(2).permute(2, 0, 1), omc["text_memory_resized"]], 0)
        add("input_proj + concat (encoder source)", rel_err(ours[0].float().view(S, B, E), o_src),
            rel_err(local[0].float().view(S, B, E), o_src))
        for i in range(len(oenc)):
            add(f"encoder layer {i} output", rel_err(ours[i + 1].float().view(S, B, E), oenc[i]),
                rel_err(local[i + 1].float().view(S, B, E), oenc[i]))
        add("img_memory (pipeline)", rel_err(mc["img_memory"], omc["img_memory"]))
        # ---- decoder: cumulative = pipeline hs; local = our decoder on the oracle's img_memory
        Lh, _, Q = oout["hs"].shape[:3]
        hs = out["_b200_hs"].float().view(Lh, Q, B, -1).transpose(1, 2)
        omc_dev = dict(mc)
        omc_dev["img_memory"] = omc["img_memory"].cuda()
        omc_dev["text_memory"] = omc_dev["img_memory"][-L:]
        out_loc = model.decode(omc_dev, want_hs=True)
        hs_loc = out_loc["_b200_hs"].float().view(Lh, Q, B, -1).transpose(1, 2)
        for l in range(Lh):
            add(f"decoder layer {l} output (after shared norm)", rel_err(hs[l], oout["hs"][l]), rel_err(hs_loc[l], oout["hs"][l]))
        # ---- heads: local = our heads on the oracle's hs
        o_hs16 = oout["hs"].cuda().transpose(1, 2).contiguous().view(Lh, Q * B, E).to(BF)
        res = R.heads_fwd(rt.call("heads", False, B=B), o_hs16, omc_dev["text_memory"])[0]
        olayers = list(oout["aux_outputs"]) + [oout]
        st = out["_b200_stacked"]
        for i, k in enumerate(("pred_logits", "pred_boxes", "proj_queries")):
            add(k + " (all layers)", rel_err(st[k], torch.stack([o[k] for o in olayers])),
                rel_err(res[i], torch.stack([o[k] for o in olayers])))
        add("proj_tokens", rel_err(st["proj_tokens"], oout["proj_tokens"]), rel_err(res[3], oout["proj_tokens"]))
    torch.cuda.synchronize()
    lines = [f"# forward error budget vs the fp32 oracle: {a.backbone}, {a.batch} x 3 x {a.size}^2, {a.tokens} tokens, "
             f"seed-0 weights", f"# {'checkpoint':58s} cumulative   stage-local"]
    for name, cum, loc in rows:
        lines.append(f"  {name:58s} {cum:.3e}    {'' if loc is None else f'{loc:.3e}'}")
    text = "\n".join(lines)
    print(text)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
