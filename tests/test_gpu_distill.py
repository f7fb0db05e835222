"""Noun-pronoun distillation on the GPU (BASELINE config 5): the warp-parallel batched LSAP,
the soft-KD loss, the distillation branch of SetCriterion and ClusterCriterion, against golden vectors frozen from the
unmodified reference (tools/make_golden.py) and against the oracle.

Bit-exact: LSAP assignments on identical cost matrices (ties, rectangular and infeasible cases included), memory-bank
bookkeeping (`update_count`, `full_label`).  Floating point: fp32 kernels, tolerances written at each assert."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import max_err, rel_err
from toist_b200.synth import make_args, make_batch, targets_to

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
DEV = "cuda"


@pytest.fixture(scope="module")
def kd_gold():
    return torch.load(GOLD / "config5_softkd_small.pt", weights_only=False)


def test_lsap_batched_is_bit_identical_to_scipy_rules():
    """One warp per problem must pick exactly what the sequential shortest-augmenting-path scan picks: compare with
    the host solver (itself pinned to scipy in tests/test_golden_cpu.py) on random, tied and rectangular problems."""
    from toist_b200 import kernels as K

    rng = np.random.RandomState(5)
    shapes = [(99, 99), (100, 100), (128, 128), (16, 16), (1, 1), (7, 128), (128, 7), (40, 64), (64, 40), (3, 5)]
    mats, dims = [], []
    for (r, c) in shapes:
        for trial in range(4):
            m = rng.rand(r, c).astype(np.float32)
            if trial == 1:
                m = (np.round(m * 3) / 3).astype(np.float32)      # many exact ties
            if trial == 2:
                m = np.zeros((r, c), np.float32)                  # all ties
            if trial == 3:
                m[:, : c // 2] = m[:, :1]                         # identical columns
            mats.append(m)
            dims.append((r, c))
    P = len(mats)
    cost = torch.zeros((P, 128, 128), dtype=torch.float32)
    for p, m in enumerate(mats):
        cost[p, : m.shape[0], : m.shape[1]] = torch.from_numpy(m)
    nr = torch.tensor([d[0] for d in dims], dtype=torch.int32, device=DEV)
    nc = torch.tensor([d[1] for d in dims], dtype=torch.int32, device=DEV)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    col = K.lsap_batched(cost.to(DEV), nr, nc, flags).cpu().numpy()
    assert int(flags.item()) == 0
    try:
        from scipy.optimize import linear_sum_assignment as solve
    except ImportError:  # pragma: no cover
        solve = K.lsap_host
    for p, m in enumerate(mats):
        r_idx, c_idx = solve(m.astype(np.float64))
        want = np.full(128, -1, np.int32)
        want[r_idx] = c_idx
        assert col[p].tolist() == want.tolist(), (dims[p], p % 4)
    # NaN and infeasible problems raise the flag instead of an exception (scipy: ValueError)
    bad = torch.zeros((2, 4, 4), dtype=torch.float32)
    bad[0, 1, 2] = float("nan")
    bad[1] = float("inf")
    four = torch.full((2,), 4, dtype=torch.int32, device=DEV)
    for p in range(2):
        flags.zero_()
        K.lsap_batched(bad[p: p + 1].contiguous().to(DEV), four[:1], four[:1], flags)
        assert int(flags.item()) == 1


def _match_from_indices(idx, t_max: int) -> torch.Tensor:
    mq = torch.full((1, len(idx), t_max), -1, dtype=torch.int32)
    for b, (src, tgt) in enumerate(idx):
        mq[0, b, tgt] = src.int()
    return mq


def test_softkd_cases_match_reference_golden(kd_gold):
    """loss_softkd + softkd_matcher (models/mdetr.py:520-599) on hand-made predictions: value and gradient."""
    from toist_b200 import kernels as K

    for c in kd_gold["softkd_cases"]:
        counts = [len(b) for b in c["tgt_boxes"]]
        t_max = max(max(counts), 1)
        mq_n = _match_from_indices(c["idx_noun"], t_max).to(DEV)
        mq_s = _match_from_indices(c["idx_sth"], t_max).to(DEV)
        cnt = torch.tensor(counts, dtype=torch.int32, device=DEV)
        flags = torch.zeros(1, dtype=torch.int32, device=DEV)
        ln, ls = c["logits_noun"][None].contiguous().to(DEV), c["logits_sth"][None].contiguous().to(DEV)
        bn, bs = c["boxes_noun"][None].contiguous().to(DEV), c["boxes_sth"][None].contiguous().to(DEV)
        loss, ws = K.softkd_fwd(ln, ls, bn, bs, mq_n, mq_s, cnt, flags)
        assert int(flags.item()) == 0
        assert abs(float(loss[0]) - c["loss"]) <= 2e-5 * max(1.0, abs(c["loss"])), (float(loss[0]), c["loss"])
        bi_n, bi_s, pair, n_fp = ws[:4]
        d = K.softkd_bwd(ls, bi_n, bi_s, pair, cnt, n_fp, torch.ones(1, device=DEV), t_max)
        assert rel_err(d[0], c["grad_logits_sth"]) < 1e-4
        # every student query is paired exactly once, with distinct teacher queries
        pr = pair[0].cpu()
        for b in range(pr.shape[0]):
            assert sorted(pr[b].tolist()) == list(range(pr.shape[1]))


def test_distillation_criterion_on_reference_outputs(kd_gold):
    """The whole list branch of SetCriterion.forward (models/mdetr.py:887-989) fed with the reference's own fp32
    predictions of teacher and student: all 66 terms."""
    from toist_b200.models import build_model

    g = kd_gold
    args = make_args("resnet50", distillation=True, softkd_loss=True, softkd_coef=50.0)
    torch.manual_seed(0)
    model, criterion, _, wd = build_model(args)
    assert dict(wd) == g["weight_dict"]
    tok = model.transformer.tokenizer
    outs, tgts, pms = [], [], []
    for tag in ("noun", "sth"):
        b = g["batch_" + tag]
        _, _, captions, targets, pm = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"])
        st = {k: v.to(DEV) for k, v in g[tag].items()}
        L = st["pred_logits"].shape[0]
        ptok = g[tag + "_proj_tokens"].to(DEV)
        tokenized = tok(captions)
        o = {"pred_logits": st["pred_logits"][-1], "pred_boxes": st["pred_boxes"][-1],
             "proj_queries": st["proj_queries"][-1], "proj_tokens": ptok, "tokenized": tokenized,
             "aux_outputs": [{"pred_logits": st["pred_logits"][l], "pred_boxes": st["pred_boxes"][l],
                              "proj_queries": st["proj_queries"][l], "proj_tokens": ptok, "tokenized": tokenized}
                             for l in range(L - 1)]}
        outs.append(o)
        tgts.append(targets_to(targets, DEV))
        pms.append(pm.to(DEV))
    with torch.no_grad():
        losses = criterion([None, None], outs, tgts, pms, None)
    assert set(losses) == set(g["losses"]) and len(losses) == 66
    for k, v in g["losses"].items():
        tol = 1e-4 * max(1.0, abs(v)) if "softkd" not in k else 2e-7 + 2e-3 * abs(v)
        assert abs(float(losses[k]) - v) <= tol, (k, float(losses[k]), v)


def test_softkd_gradient_reaches_only_the_student(kd_gold):
    from toist_b200.models import build_model

    g = kd_gold
    args = make_args("resnet50", distillation=True, softkd_loss=True, softkd_coef=50.0)
    torch.manual_seed(0)
    _, criterion, _, wd = build_model(args)
    outs, tgts, pms, leaves = [], [], [], []
    for tag in ("noun", "sth"):
        b = g["batch_" + tag]
        _, _, captions, targets, pm = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"])
        lg = g[tag]["pred_logits"].to(DEV).requires_grad_(True)
        bx = g[tag]["pred_boxes"].to(DEV)
        leaves.append(lg)
        criterion.losses = ["labels", "boxes", "cardinality", "softkd"]
        L = lg.shape[0]
        outs.append({"pred_logits": lg[-1], "pred_boxes": bx[-1],
                     "aux_outputs": [{"pred_logits": lg[l], "pred_boxes": bx[l]} for l in range(L - 1)],
                     "_b200_stacked": {"pred_logits": lg, "pred_boxes": bx}})
        tgts.append(targets_to(targets, DEV))
        pms.append(pm.to(DEV))
    losses = criterion([None, None], outs, tgts, pms, None)
    kd = sum(v for k, v in losses.items() if k.startswith("loss_softkd"))
    kd.backward()
    assert leaves[0].grad is None or float(leaves[0].grad.abs().max()) == 0.0  # the teacher is detached (mdetr.py:552)
    # teacher == student at initialisation: the two-class probabilities agree to ~1e-4 and the gradient is a difference
    # of nearly equal numbers in both implementations (1e-4 accuracy on well separated inputs: the test above)
    assert rel_err(leaves[1].grad, g["softkd_grad_sth_logits"]) < 5e-2


def test_cluster_criterion_reproduces_reference_run():
    """ClusterCriterion driven exactly like the recorded reference run (10 steps, memory_size 8 so that banks fill and
    the LSAP replacement branch runs): replaced memories, loss, its gradient and every buffer after every step."""
    from toist_b200.models.cluster import ClusterCriterion
    from toist_b200.tokenizer import CharTokenizer

    g = torch.load(GOLD / "cluster_cases.pt", weights_only=False)
    d = g["dims"]
    args = make_args("resnet50", cluster=True, cluster_memory_size=d["memory_size"], cluster_num=d["cluster_num"],
                     train_batch_size=d["B"])
    cc = ClusterCriterion(feature_dim=d["D"], memory_size=d["memory_size"], cluster_num=d["cluster_num"], task_count=14,
                          args=args).to(DEV)
    cc.load_state_dict(g["init"])
    np.random.seed(g["numpy_seed"])
    tok = CharTokenizer()
    for step, rec in enumerate(g["steps"]):
        B = d["B"]
        tg_n = [{"boxes": torch.zeros(rec["nbox"][i], 4, device=DEV), "dataset_name": f"tdod_{rec['tasks'][i]}",
                 "noun_tokens_positive": rec["noun_tokens_positive"][i]} for i in range(B)]
        tg_s = [{"boxes": torch.zeros(1, 4, device=DEV), "dataset_name": f"tdod_{rec['tasks'][i]}"} for i in range(B)]
        # teacher side
        img = rec["img_n"].to(DEV)
        mc = {"img_memory": img, "tokenized": tok(rec["cap_n"])}
        T = mc["tokenized"]["input_ids"].shape[1]
        mc["text_memory"] = mc["img_memory"][-T:]
        mc = cc.update_memory(mc, tg_n, rec["cap_n"])
        assert rel_err(mc["img_memory_mod"], rec["mod_n"]) < 1e-5, step
        # student side
        img = rec["img_s"].to(DEV).requires_grad_(True)
        mc = {"img_memory": img, "tokenized": tok(rec["cap_s"])}
        T = mc["tokenized"]["input_ids"].shape[1]
        mc["text_memory"] = mc["img_memory"][-T:]
        mc, loss = cc(mc, tg_s, rec["cap_s"])
        assert rel_err(mc["img_memory_mod"], rec["mod_s"]) < 1e-5, step
        assert float(loss["loss_cluster_choice"]) == 0.0
        v = rec["loss"]["loss_cluster_feature"]
        assert abs(float(loss["loss_cluster_feature"]) - v) <= 1e-5 * max(1.0, abs(v)), step
        (loss["loss_cluster_feature"] * 3.0 + mc["img_memory_mod"].sum()).backward()
        assert rel_err(img.grad, rec["grad_img_s"]) < 1e-5, step
        sd = cc.state_dict()
        assert torch.equal(sd["update_count"].cpu(), rec["state"]["update_count"]), step
        assert torch.equal(sd["full_label"].cpu(), rec["state"]["full_label"]), step
        assert max_err(sd["feature_bank"], rec["state"]["feature_bank"]) < 1e-5, step
        assert max_err(sd["cluster_centers"], rec["state"]["cluster_centers"]) < 1e-5, step


def test_kmeans_kernel_against_oracle():
    from oracle import model as O
    from toist_b200 import kernels as K

    torch.manual_seed(4)
    for N, D, Kc in ((1024, 256, 3), (64, 32, 5), (10, 8, 3)):
        X = torch.randn(N, D)
        X[: N // 3] += 2.0
        init = X[torch.randperm(N)[:Kc]].clone()
        choice_o, centers_o = O.kmeans(X, init.clone(), Kc, full_label=1.0)
        centers = init.clone().to(DEV)
        choice, iters = K.kmeans(X.to(DEV), centers)
        assert int(iters.item()) >= 1
        assert max_err(centers, centers_o) < 1e-5
        assert choice.cpu().tolist() == choice_o.tolist()
        q = torch.randn(7, D)
        assert K.kmeans_predict(q.to(DEV), centers).cpu().tolist() == O.kmeans_predict(q, centers_o).tolist()


def test_distillation_step_end_to_end():
    """engine.py:182-204 with two models, ClusterCriterion and soft-KD on the GPU: runs, every loss is finite, both
    models and the text branch of the student receive gradients; the noun_/sth_ halves equal the single-model
    criterion on the same predictions; soft-KD agrees with the oracle on our own predictions."""
    from copy import deepcopy

    from oracle import model as O
    from toist_b200.models import build_model
    from toist_b200.util.misc import NestedTensor

    args = make_args("resnet50", distillation=True, softkd_loss=True, softkd_coef=50.0, cluster=True,
                     cluster_memory_size=16, cluster_num=3, train_batch_size=2)
    torch.manual_seed(0)
    np.random.seed(0)
    model, criterion, cluster_criterion, wd = build_model(args)
    assert {"loss_softkd", "loss_cluster_feature", "noun_loss_ce", "sth_loss_giou_4", "loss_softkd_4"} <= set(wd)
    model.cuda().eval()
    model_noun = deepcopy(model)
    cluster_criterion.cuda()
    cluster_criterion.syn_memory()
    bn = make_batch(2, 128, 16, seed=31, pad=True)   # 16 tokens: the caption has room for the word 'something'
    bs = make_batch(2, 128, 16, seed=32, pad=False)
    for t in bn[3] + bs[3]:
        t["dataset_name"] = "tdod_3"
    tn, ts = targets_to(bn[3], DEV), targets_to(bs[3], DEV)
    sn = NestedTensor(bn[0].cuda(), bn[1].cuda())
    ss = NestedTensor(bs[0].cuda(), bs[1].cuda())
    mc_n = model_noun(sn, bn[2], encode_and_save=True)
    mc_n = cluster_criterion.update_memory(mc_n, tn, bn[2])
    out_n = model_noun(sn, bn[2], encode_and_save=False, memory_cache=mc_n)
    mc_s = model(ss, bs[2], encode_and_save=True)
    mc_s, loss_cluster = cluster_criterion(mc_s, ts, bs[2])
    out_s = model(ss, bs[2], encode_and_save=False, memory_cache=mc_s)
    losses = criterion([mc_n, mc_s], [out_n, out_s], [tn, ts], [bn[4].cuda(), bs[4].cuda()], None)
    losses.update(loss_cluster)
    assert len(losses) == 68
    total = sum(losses[k] * wd[k] for k in losses if k in wd)
    assert torch.isfinite(total)
    total.backward()
    for m in (model, model_noun):
        assert m.transformer.decoder.layers[0].linear1.weight.grad is not None
        assert m.backbone[0].body.layer3[0].conv1.weight.grad is not None
    assert model.transformer.text_encoder.embeddings.word_embeddings.weight.grad is not None
    # halves == single-model criterion on the same predictions
    with torch.no_grad():
        single = criterion(mc_s, out_s, ts, bs[4].cuda(), None)
    for k, v in single.items():
        assert abs(float(v) - float(losses["sth_" + k])) <= 1e-6 * max(1.0, abs(float(v))), k
    # soft-KD of the last layer against the oracle on our fp32 predictions and our assignments
    (mq_n, cnt_n, _), (mq_s, cnt_s, _) = criterion.last_match_pair
    from toist_b200.models.matcher import indices_from_match

    idx_n = indices_from_match(mq_n[-1].cpu(), cnt_n)
    idx_s = indices_from_match(mq_s[-1].cpu(), cnt_s)
    on = {"pred_logits": out_n["pred_logits"].detach().cpu(), "pred_boxes": out_n["pred_boxes"].detach().cpu()}
    os_ = {"pred_logits": out_s["pred_logits"].detach().cpu(), "pred_boxes": out_s["pred_boxes"].detach().cpu()}
    want = float(O.loss_softkd(on, os_, idx_n, idx_s, 100))
    assert abs(float(losses["loss_softkd"]) - want) <= 2e-7 + 2e-3 * abs(want)


def test_distillation_criterion_graph_replay_matches_eager(kd_gold):
    """SetCriterion evaluates its stage twice per distillation step (noun, then sth) with identical shapes and flags:
    each pass must own its captured graph, otherwise the second replay overwrites the first pass's losses, assignments
    and saved gradients.  Eager vs captured vs replayed, values and gradients."""
    from toist_b200.models import build_model
    from toist_b200.tokenizer import CharTokenizer

    g = kd_gold
    args = make_args("resnet50", distillation=True, softkd_loss=True, softkd_coef=50.0)
    _, criterion, _, wd = build_model(args)
    tok = CharTokenizer()
    def mk(b):
        return make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"])

    batches = {"noun": mk(g["batch_noun"]), "sth": mk(g["batch_sth"])}

    def run():
        outs, leaves, tg, pms = [], [], [], []
        for tag in ("noun", "sth"):
            _, _, captions, targets, pm = batches[tag]
            tokd = tok(captions)
            L = g[tag]["pred_logits"].shape[0]
            st = {k: g[tag][k].to(DEV).clone().requires_grad_(True) for k in ("pred_logits", "pred_boxes", "proj_queries")}
            ptok = g[tag + "_proj_tokens"].to(DEV).clone().requires_grad_(True)
            layers = [{**{k: st[k][l] for k in st}, "proj_tokens": ptok, "tokenized": tokd} for l in range(L)]
            o = dict(layers[-1])
            o["aux_outputs"] = layers[:-1]
            outs.append(o)
            leaves.append((st, ptok))
            tg.append(targets_to(targets, DEV))
            pms.append(pm.to(DEV))
        losses = criterion([{}, {}], outs, tg, pms, None)
        total = sum(losses[k] * wd[k] for k in losses if k in wd)
        total.backward()
        vals = {k: float(v) for k, v in losses.items()}
        grads = [t.grad.clone() for st, ptok in leaves for t in (*st.values(), ptok)]
        return vals, grads

    v0, g0 = run()
    for k, v in g["losses"].items():  # the eager run itself reproduces the reference's 66 terms
        assert abs(v0[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, v0[k], v)
    assert abs(v0["noun_loss_ce"] - v0["sth_loss_ce"]) > 1e-6  # the two halves differ: aliasing would be visible
    criterion.enable_cuda_graphs(True)
    for _ in range(3):  # capture, replay, replay
        v1, g1 = run()
        assert v1 == v0
        for a, b in zip(g0, g1):
            assert torch.equal(a, b)
    criterion.enable_cuda_graphs(False)


def test_loss_nsthl2_matches_reference_golden():
    """SetCriterion's noun / pronoun text-feature L2 term (models/mdetr.py:668-781) on the teacher / student text
    memories frozen from the reference (tests/golden/nsthl2_cases.pt): loss value and d loss / d student text memory,
    incl. an image without targets, an image whose teacher has no boxes, and the all-empty batch (constant zero)."""
    from toist_b200.models.mdetr import SetCriterion
    from toist_b200.tokenizer import CharTokenizer

    gold = torch.load(GOLD / "nsthl2_cases.pt", weights_only=False)
    args = make_args("resnet50", distillation=True, nsthl2_loss=True)
    crit = SetCriterion(args, 255, matcher=None, eos_coef=0.1, losses=["labels", "boxes", "cardinality", "nsthl2"],
                        temperature=0.07, contrastive_hdim=64)
    tok = CharTokenizer()
    for case in gold["cases"]:
        mcs, outs, tgts = [], [], []
        for k, key in enumerate(("text_noun", "text_sth")):
            text = case[key].to(DEV).requires_grad_(k == 1)
            mcs.append({"text_memory": text})
            outs.append({"tokenized": tok.batch_encode_plus(case["captions"][k], padding="longest", return_tensors="pt")})
            tgts.append([{"noun_tokens_positive": spans} for spans in case["noun_tokens_positive"][k]])
        val = crit._loss_nsthl2(mcs, outs, tgts, case["counts"][1])
        assert abs(float(val) - case["loss"]) <= 1e-6 + 1e-5 * abs(case["loss"]), (float(val), case["loss"])
        if case["grad_text_sth"] is None:
            assert not val.requires_grad and float(val) == 0.0
            continue
        val.backward()
        assert mcs[0]["text_memory"].grad is None  # the teacher's features are detached (:775)
        assert rel_err(mcs[1]["text_memory"].grad.cpu(), case["grad_text_sth"]) <= 1e-5
    # the weight dict carries the term for the last layer and every auxiliary layer, without model prefix
    from toist_b200.models import build_model

    wd = build_model(make_args("resnet50", distillation=True, nsthl2_loss=True))[3]
    assert sorted(k for k in wd if "nsthl2" in k) == sorted(gold["weight_keys"])
