"""ClusterCriterion: per-task memory bank of noun features, k-means prototypes and the pronoun-feature replacement of
the distillation recipe (BASELINE config 5), behind the reference's interface (models/mdetr.py:29-312).

Same constructor, buffers (`feature_bank [14, M, D]`, `cluster_centers [14, K, D]`, `update_count [14]`,
`full_label [14]`), methods and return values as the reference.  What changes is where the arithmetic runs:

* token means, the replacement of the caption tokens, the cluster-feature MSE and their gradients are kernels of
  csrc/distill.cu; the Lloyd iterations of models/kmeans.py run in ONE launch per call instead of one host
  synchronisation per iteration (`toist_kmeans`);
* `full_label` / `update_count` are mirrored on the host (they only depend on how many features arrived), so the
  reference's `.item()` round trips disappear; the buffers stay authoritative for `state_dict()`.

The random k-means initialisation keeps the reference's quirk: rows are drawn with numpy's global RNG
(`models/kmeans.py:16`), seeded per rank in `main.py:308-310`.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
from torch import nn

from .. import kernels as K
from ..util import dist
from ..util.misc import h2d
from .mdetr import _token_spans


class _TokenWeightedSum(torch.autograd.Function):
    """feat[b] = sum_t w[b, t] * text[t, b]  (the token means of models/mdetr.py:141,256), differentiable in `text`."""

    @staticmethod
    def forward(ctx, text, w):
        ctx.save_for_backward(w)
        ctx.T = text.shape[0]
        return K.token_wsum(text.contiguous(), w)

    @staticmethod
    def backward(ctx, dout):
        (w,) = ctx.saved_tensors
        return K.token_wsum_bwd(dout.contiguous(), w, ctx.T), None


class _TokenReplace(torch.autograd.Function):
    """out = img_memory.clone(); out[-T:, b][selected tokens] = feat[b]  (models/mdetr.py:186,208,243,266).  The
    replaced rows receive no gradient, every other row passes it through."""

    @staticmethod
    def forward(ctx, img_memory, sel, feat, n_text: int):
        out = img_memory.clone()
        if n_text:
            K.token_fill(out[-n_text:], sel, feat)
        ctx.save_for_backward(sel)
        ctx.n_text = n_text
        return out

    @staticmethod
    def backward(ctx, dout):
        (sel,) = ctx.saved_tensors
        d = dout.clone()
        if ctx.n_text:
            K.token_fill(d[-ctx.n_text:], sel, None)
        return d, None, None, None


class _MseRows(torch.autograd.Function):
    """mean over the used rows of F.mse_loss(a[r], b[r])  (loss_cluster_feature, models/mdetr.py:270-278)."""

    @staticmethod
    def forward(ctx, a, b, use):
        loss, da = K.mse_rows(a.contiguous(), b.contiguous(), use, True)
        ctx.save_for_backward(da)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (da,) = ctx.saved_tensors
        return da * g, None, None


def _selection(tokenized, i: int, spans_per_box, n_tokens: int) -> List[np.ndarray]:
    """One 0/1 vector over the caption tokens per box (models/mdetr.py:123-150)."""
    out = []
    for spans in spans_per_box:
        pos = np.zeros(n_tokens, dtype=np.float32)
        for beg, end in _token_spans(tokenized, i, spans):
            pos[beg: end + 1] = 1
        out.append(pos)
    return out


class ClusterCriterion(nn.Module):
    def __init__(self, feature_dim, memory_size, cluster_num, task_count, args):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("toist_b200 ClusterCriterion runs on CUDA devices only; there is no CPU path")
        self.args = args
        self.feature_dim = feature_dim
        self.memory_size = memory_size
        self.cluster_num = cluster_num
        self.task_count = task_count
        self.temp_feature_idx_list = [torch.zeros([args.train_batch_size, feature_dim + 1]).cuda()
                                      for _ in range(dist.get_world_size())]
        self.register_buffer("feature_bank", torch.randn([task_count, memory_size, feature_dim]))
        self.register_buffer("cluster_centers", torch.randn([task_count, cluster_num, feature_dim]))
        self.register_buffer("update_count", torch.zeros([task_count]))
        self.register_buffer("full_label", torch.zeros([task_count]))
        self._host_state_stale = True
        self._h_full: List[float] = []
        self._h_count: List[float] = []

    # ---- host mirror of full_label / update_count
    def _apply(self, fn, *a, **k):
        self._host_state_stale = True
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._host_state_stale = True
        return super()._load_from_state_dict(*a, **k)

    def _host_state(self):
        if self._host_state_stale:
            self._h_full = [float(v) for v in self.full_label.tolist()]
            self._h_count = [float(v) for v in self.update_count.tolist()]
            self._host_state_stale = False
        return self._h_full, self._h_count

    def syn_memory(self):
        """models/mdetr.py:52-60: average bank and centres over the ranks so that every rank starts identically."""
        world_size = dist.get_world_size()
        if dist.is_dist_avail_and_initialized():
            torch.distributed.all_reduce(self.feature_bank)
            torch.distributed.all_reduce(self.cluster_centers)
        self.feature_bank /= world_size
        self.cluster_centers /= world_size

    # ---- memory bank
    def update_memory_queue(self, feature_idx_list: torch.Tensor, task_idx_host: Optional[List[int]] = None):
        """models/mdetr.py:62-103.  `feature_idx_list` [B, D + 1]: feature and task index (-1 = empty) per sample.
        `task_idx_host` is the same last column when the caller already knows it on the host (single rank)."""
        if dist.is_dist_avail_and_initialized():
            for t in self.temp_feature_idx_list:
                t.zero_()
            torch.distributed.all_gather(self.temp_feature_idx_list, feature_idx_list)
            gathered = torch.cat(self.temp_feature_idx_list, dim=0)
            tasks = None if dist.get_world_size() > 1 else task_idx_host
        else:
            gathered, tasks = feature_idx_list, task_idx_host
        if tasks is None:
            tasks = [int(v) for v in gathered[:, -1].tolist()]  # other ranks' task ids: one small device read
        full, count = self._host_state()
        rows_of = {}
        for r, t in enumerate(tasks):
            if t != -1:
                rows_of.setdefault(int(t), []).append(r)
        for i in sorted(rows_of):
            rows = rows_of[i]
            n = len(rows)
            new = gathered[torch.as_tensor(rows, device=gathered.device), :-1]
            bank = self.feature_bank[i]
            if full[i] == 0 or self.args.fifo_memory:
                bank[:-n] = bank[n:].clone()
                bank[-n:] = new
                if full[i] == 0:
                    if count[i] > self.memory_size:
                        full[i] = 1.0
                        self.full_label[i] = 1
                    count[i] += n
                    self.update_count[i] += n
            else:  # replace the nearest stored features (L1), one per new feature
                l1 = K.cdist_l1(new.contiguous(), bank.contiguous())
                r_idx, c_idx = K.lsap_host(l1.cpu().numpy())
                bank[torch.as_tensor(c_idx, device=bank.device)] = new[torch.as_tensor(r_idx, device=bank.device)]

    def memory_cluster(self, feature_to_cluster: torch.Tensor, task_idx: int):
        """models/mdetr.py:213-234: k-means on the task's bank (random rows as initial centres until the bank is full,
        the stored centres afterwards), then the nearest centre of `feature_to_cluster` [D].
        Returns (choice int32 [1] on the device, centre feature [D])."""
        full, _ = self._host_state()
        bank = self.feature_bank[task_idx]
        if full[task_idx] == 0:
            pick = np.random.choice(bank.shape[0], self.cluster_num, replace=False)
            centers = bank[torch.as_tensor(pick, device=bank.device)].contiguous()
        else:
            centers = self.cluster_centers[task_idx].clone()
        K.kmeans(bank, centers, tol=1e-4)
        self.cluster_centers[task_idx] = centers
        choice = K.kmeans_predict(feature_to_cluster.reshape(1, -1).contiguous(), centers)
        feature = centers.index_select(0, choice.long())[0]
        return choice, feature

    def memory_cluster_many(self, feats: torch.Tensor, tasks: List[int], active: List[bool]) -> torch.Tensor:
        """`memory_cluster` for every active sample of a batch, in the reference's order (models/mdetr.py:196-209,
        260-270 loop over the samples), with the samples whose tasks are pairwise distinct clustered by ONE launch:
        k-means runs on different banks are independent; a sample whose task already occurred in the current group
        starts a new group, so that it sees the centres its predecessor left (the sequential dependency of the loop).
        numpy's generator is drawn from in sample order, as the loop does.  Returns [bs, D]: the centre chosen for each
        sample, read AFTER its group's update of `cluster_centers` (zeros for inactive samples)."""
        bs, D = feats.shape
        dev = feats.device
        full, _ = self._host_state()
        chosen = torch.zeros((bs, D), dtype=torch.float32, device=dev)
        group: List[int] = []
        inits: List[torch.Tensor] = []

        def flush():
            if not group:
                return
            tk = torch.as_tensor([tasks[i] for i in group], dtype=torch.int64)
            task_dev = h2d(tk.to(torch.int32), dev)
            centers = torch.stack(inits).contiguous()
            rows = h2d(torch.as_tensor(group, dtype=torch.int64), dev)
            q = feats.index_select(0, rows).contiguous()
            qc = K.kmeans_batched(self.feature_bank, task_dev, centers, q)
            task_long = h2d(tk, dev)
            self.cluster_centers.index_copy_(0, task_long, centers)
            pick = torch.arange(len(group), device=dev) * self.cluster_num + qc.long()
            chosen.index_copy_(0, rows, centers.view(-1, D).index_select(0, pick))
            group.clear()
            inits.clear()

        for i in range(bs):
            if not active[i]:
                continue
            t = tasks[i]
            if any(tasks[j] == t for j in group):
                flush()
            bank = self.feature_bank[t]
            if full[t] == 0:
                pick = np.random.choice(bank.shape[0], self.cluster_num, replace=False)
                inits.append(bank[h2d(torch.as_tensor(pick, dtype=torch.int64), dev)])
            else:
                # (a task repeated inside the batch was flushed above: the stored centres are up to date)
                inits.append(self.cluster_centers[t].clone())
            group.append(i)
        flush()
        return chosen

    # ---- teacher side
    def update_memory(self, memory_cache_noun, targets_noun, captions_noun):
        """models/mdetr.py:105-211."""
        text = memory_cache_noun["text_memory"]  # [T, B, D]
        T, bs, D = text.shape
        dev = text.device
        tokenized = memory_cache_noun["tokenized"]
        w = np.zeros((bs, T), dtype=np.float32)     # mean over the boxes of the per-box token means
        sel = np.zeros((bs, T), dtype=np.uint8)     # union of the noun tokens of all boxes
        skip = [len(t["boxes"]) == 0 for t in targets_noun]
        for i, tgt in enumerate(targets_noun):
            per_box = _selection(tokenized, i, tgt["noun_tokens_positive"], T)
            for pos in per_box:
                n = pos.sum()
                w[i] += (pos / n if n > 0 else np.full(T, np.nan, np.float32)) / len(per_box)
                sel[i] = np.maximum(sel[i], pos.astype(np.uint8))
        w_dev = torch.from_numpy(w).to(dev)
        with torch.no_grad():
            feats = K.token_wsum(text.detach().contiguous(), w_dev)  # token_feature_all_noun
        tasks = [-1 if skip[i] else int(targets_noun[i]["dataset_name"].split("_")[1]) - 1 for i in range(bs)]
        keep = torch.as_tensor([0.0 if s else 1.0 for s in skip], device=dev).view(-1, 1)
        feature_list = torch.where(keep.bool(), feats, torch.zeros_like(feats))
        task_col = torch.as_tensor(tasks, dtype=torch.float32, device=dev).view(-1, 1)
        with torch.no_grad():
            self.update_memory_queue(torch.cat([feature_list, task_col], dim=-1), tasks)
            for i in range(bs):
                if skip[i]:
                    sel[i] = 0
            chosen = self.memory_cluster_many(feats, tasks, [not s_ for s_ in skip])
        sel_dev = torch.from_numpy(sel).to(dev)
        memory_cache_noun["img_memory_mod"] = _TokenReplace.apply(memory_cache_noun["img_memory"], sel_dev, chosen, T)
        memory_cache_noun["full_label"] = self.full_label
        memory_cache_noun["update_count"] = self.update_count
        return memory_cache_noun

    # ---- student side
    def _something(self, tokenized, captions, T: int):
        w = np.zeros((len(captions), T), dtype=np.float32)
        sel = np.zeros((len(captions), T), dtype=np.uint8)
        for i, cap in enumerate(captions):
            begin = cap.find("something")
            end = begin + len("something")
            beg_pos = tokenized.char_to_token(i, begin)
            end_pos = tokenized.char_to_token(i, end - 1)
            sel[i, beg_pos: end_pos + 1] = 1
            w[i] = sel[i] / max(int(sel[i].sum()), 1) if sel[i].sum() else np.nan
        return w, sel

    def forward(self, memory_cache_sth, targets_sth, captions_sth):
        """models/mdetr.py:236-280: returns (memory_cache with `img_memory_mod`, {loss_cluster_choice, loss_cluster_feature})."""
        text = memory_cache_sth["text_memory"]
        T, bs, D = text.shape
        dev = text.device
        w, sel = self._something(memory_cache_sth["tokenized"], captions_sth, T)
        w_dev, sel_dev = torch.from_numpy(w).to(dev), torch.from_numpy(sel).to(dev)
        feats = _TokenWeightedSum.apply(text, w_dev)  # temp_token_feature of every sample, differentiable
        tasks = [int(t["dataset_name"].split("_")[1]) - 1 for t in targets_sth]
        with torch.no_grad():
            # cluster_centers[task, choice] right after the sample's own k-means: both the replacement feature and the
            # centre the loss compares against (models/mdetr.py:262-270 read the same tensor)
            chosen = self.memory_cluster_many(feats.detach().contiguous(), tasks, [True] * bs)
            centre = chosen
        memory_cache_sth["img_memory_mod"] = _TokenReplace.apply(memory_cache_sth["img_memory"], sel_dev, chosen, T)
        use = torch.ones(bs, dtype=torch.uint8, device=dev)
        loss_feature = _MseRows.apply(feats, centre, use) if bs else torch.zeros((), device=dev)
        loss_choice = torch.zeros((), device=dev)
        return memory_cache_sth, {"loss_cluster_choice": loss_choice, "loss_cluster_feature": loss_feature}

    def infer_choice(self, memory_cache_sth, dataset_name_list, captions):
        """models/mdetr.py:282-312 (evaluation: replacement only, no loss)."""
        text = memory_cache_sth["text_memory"]
        T, bs, D = text.shape
        dev = text.device
        w, sel = self._something(memory_cache_sth["tokenized"], captions, T)
        with torch.no_grad():
            feats = K.token_wsum(text.detach().contiguous(), torch.from_numpy(w).to(dev))
            tasks = [int(name.split("_")[1]) - 1 for name in dataset_name_list]
            chosen = self.memory_cluster_many(feats, tasks, [True] * bs)
        memory_cache_sth["img_memory_mod"] = _TokenReplace.apply(memory_cache_sth["img_memory"],
                                                                 torch.from_numpy(sel).to(dev), chosen, T)
        return memory_cache_sth
