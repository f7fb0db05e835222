"""The two torch.distributed helpers the criterion needs (reference util/dist.py:150-170)."""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch
import torch.distributed as dist
from torch import nn


def is_dist_avail_and_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_world_size() -> int:
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process() -> bool:
    return get_rank() == 0


def reduce_dict(input_dict, average: bool = True):
    """util/dist.py:93-117: the loss dictionary averaged over ranks for logging (engine.py:75); keys sorted so that every
    rank stacks them in the same order; the dictionary itself when there is a single rank."""
    world = get_world_size()
    if world < 2:
        return input_dict
    with torch.no_grad():
        names = sorted(input_dict.keys())
        values = torch.stack([input_dict[k] for k in names], dim=0)
        dist.all_reduce(values)
        if average:
            values /= world
        return {k: v for k, v in zip(names, values)}


# ------------------------------------------------------------------------------------------------ gradient all-reduce
class FlatGradSync:
    """Gradient exchange of data-parallel training (reference main.py:336: DistributedDataParallel's bucketed
    all-reduce).  Every backward stage of the model (heads, decoder, encoder, backbone, text) already writes ALL of its
    parameter gradients into one flat fp32 arena (blocks.zero_arena) and `param.grad` are views of it, so the exchange
    is one in-place NCCL all-reduce (average) per stage on a side stream, issued the moment the stage's backward has
    been queued: it overlaps the backward of the stages that follow, needs no per-parameter hooks, no bucket copies and
    no autograd-graph traversal.  The calling stream waits for the side stream at the end of the backward pass."""

    def __init__(self, process_group=None):
        self.group = process_group
        self.stream: Optional[torch.cuda.Stream] = None
        self.enabled = True
        self._pending = False
        self._outside = {}
        # TOIST_GRAD_MARKS=0: one all-reduce per stage when its backward has been queued (round-1 behaviour);
        # TOIST_SPARSE_EMBED=0: the word-embedding gradient travels inside the dense arena
        self.marks = os.environ.get("TOIST_GRAD_MARKS", "1") != "0"
        self.sparse = os.environ.get("TOIST_SPARSE_EMBED", "1") != "0"
        self.sparse_rows = int(os.environ.get("TOIST_SPARSE_EMBED_ROWS", "2048"))  # (id, row) pairs per rank, fixed
        self._sparse_buf = {}
        # The text branch's backward runs NEXT TO the trunk's: its gradient chunks become final in between the trunk's.
        # One communicator executes its collectives in issue order, so the text stage gets a communicator and a stream
        # of its own (set by DistributedDataParallel, created collectively); everything else shares the default one.
        self.lanes = {}  # stage name -> (process group, stream)
        self.n_extra = 0  # gradients exchanged one by one because they were not served from a stage arena (should stay 0)
        self._used = set()

    def _lane(self, stage_name: str, device):
        lane = self.lanes.get(stage_name)
        if lane is not None:
            if lane[1] is None or lane[1].device != device:
                lane = self.lanes[stage_name] = (lane[0], torch.cuda.Stream(device=device))
            return lane
        if self.stream is None or self.stream.device != device:
            self.stream = torch.cuda.Stream(device=device)
        return self.group, self.stream

    def wants_marks(self) -> bool:
        """Stages record an event every blocks.zero_arena.mark_bytes of finished gradients (a prefix of their arena),
        so the exchange of a long backward stage (RoBERTa, the trunk) starts while the stage is still running."""
        return (self.marks and self.enabled and is_dist_avail_and_initialized() and get_world_size() > 1
                and torch.cuda.is_available())

    def reduce(self, stage_name: str, flat: torch.Tensor, grads, blocking: bool = False, marks=(), sparse=None) -> None:
        if not self.enabled or not is_dist_avail_and_initialized() or get_world_size() == 1:
            return
        if not flat.is_cuda:  # host tensors (gloo, used by the CPU tests of this logic): SUM then divide, in line
            lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
            for t in [flat] + [g for g in grads if g is not None and not (lo <= g.data_ptr() < hi)]:
                dist.all_reduce(t, group=self.group)
                t /= get_world_size()
            return
        cur = torch.cuda.current_stream(flat.device)
        group, stream = self._lane(stage_name, flat.device)
        # gradients that were not served from the arena (none today) are exchanged one by one; checked once per stage
        extra = []
        if grads:
            key = (stage_name, flat.data_ptr(), flat.numel())
            if key not in self._outside:
                if len(self._outside) > 64:  # eager mode allocates a new arena every step: do not grow without bound
                    self._outside.clear()
                lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
                self._outside[key] = [i for i, t in enumerate(grads) if t is not None and not (lo <= t.data_ptr() < hi)]
            extra = [grads[i] for i in self._outside[key]]
            self.n_extra += len(extra)
        # the word-embedding table is the tail of the text stage's arena (runtime.text_bwd): left out of the dense
        # exchange when it can travel as (id, row) pairs
        dense_end = flat.numel()
        if sparse is not None:
            _, ids, rows, pad_id, table = sparse
            off = (table.data_ptr() - flat.data_ptr()) // 4
            ok = (self.sparse and 0 <= off and off + table.numel() <= flat.numel() and ids.numel() <= self.sparse_rows
                  and all(t is None or not (table.data_ptr() < t.data_ptr() < flat.data_ptr() + flat.numel() * 4)
                          for t in grads))
            if ok:
                dense_end = off
            else:
                sparse = None
        prev = 0
        with torch.cuda.stream(stream):
            for off, ev in marks:  # prefixes of the arena that are final while the stage's backward is still running
                off = min(int(off), dense_end)
                if off <= prev:
                    continue
                stream.wait_event(ev)
                dist.all_reduce(flat[prev:off], op=dist.ReduceOp.AVG, group=group)
                prev = off
        stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            if dense_end > prev:
                dist.all_reduce(flat[prev:dense_end], op=dist.ReduceOp.AVG, group=group)
            for t in extra:
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
            if sparse is not None:
                self._sparse_exchange(ids, rows, int(pad_id), table, group)
        if blocking:
            cur.wait_stream(stream)
        else:
            self._used.add(stream)
            if not self._pending:
                self._pending = True
                torch.autograd.Variable._execution_engine.queue_callback(self._join)

    def _sparse_exchange(self, ids: torch.Tensor, rows: torch.Tensor, pad_id: int, table: torch.Tensor, group) -> None:
        """All ranks contribute a FIXED number of (token id, gradient row) pairs (unused slots carry the padding id, which
        the merge skips), gathered in rank order; every rank then writes scale * (sum per id) into its table."""
        from .. import kernels as K

        world = get_world_size()
        m, e = rows.shape
        key = (rows.device, e, world)
        buf = self._sparse_buf.get(key)
        if buf is None:
            r = self.sparse_rows
            buf = self._sparse_buf[key] = (torch.empty((r, e), dtype=torch.float32, device=rows.device),
                                           torch.empty((r,), dtype=torch.int64, device=rows.device),
                                           torch.empty((world * r, e), dtype=torch.float32, device=rows.device),
                                           torch.empty((world * r,), dtype=torch.int64, device=rows.device))
        my_rows, my_ids, all_rows, all_ids = buf
        my_ids.fill_(pad_id)
        my_ids[:m].copy_(ids)
        my_rows[:m].copy_(rows)
        dist.all_gather_into_tensor(all_ids, my_ids, group=group)
        dist.all_gather_into_tensor(all_rows, my_rows, group=group)
        K.embed_rows_merge(table, all_ids, all_rows, pad_id, 1.0 / world)

    def _join(self) -> None:
        self._pending = False
        for st in self._used:
            torch.cuda.current_stream(st.device).wait_stream(st)
        self._used.clear()


class DistributedDataParallel(nn.Module):
    """Drop-in for `torch.nn.parallel.DistributedDataParallel(model, device_ids=[gpu], find_unused_parameters=True)`
    at reference main.py:336,345 for toist_b200 models: same constructor keywords, `.module`, `no_sync()`, state dict
    with the `module.` prefix.  Parameters and buffers are broadcast from rank 0 once at construction (as torch's
    wrapper does); gradients are averaged by FlatGradSync.  Parameters that receive no gradient in a step (RoBERTa's
    pooler) simply keep `grad = None` on every rank: there is nothing to find or to wait for."""

    def __init__(self, module: nn.Module, device_ids=None, output_device=None, find_unused_parameters: bool = False,
                 process_group=None, broadcast_buffers: bool = True, **_unused):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.grad_sync = FlatGradSync(process_group)
        inner = getattr(module, "detr", module)  # DETRsegm wraps the detector (models/segmentation.py:29)
        rt = getattr(inner, "_rt", None)
        if rt is None or not hasattr(rt, "grad_sync"):
            raise TypeError("toist_b200.util.dist.DistributedDataParallel wraps toist_b200 models (MDETR / DETRsegm)")
        rt.grad_sync = self.grad_sync
        rt.direct = True  # stages assign .grad themselves: nothing here listens on AccumulateGrad (runtime.StageFn)
        if (is_dist_avail_and_initialized() and get_world_size() > 1 and dist.get_backend(process_group) == "nccl"
                and os.environ.get("TOIST_TEXT_COMM", "1") != "0"):
            ranks = dist.get_process_group_ranks(process_group) if process_group is not None else None
            self.grad_sync.lanes["text"] = (dist.new_group(ranks=ranks, backend="nccl"), None)  # collective: all ranks
        covered = ()
        if inner is not module:  # the mask branch is a stage of its own
            module._rt.grad_sync = self.grad_sync
            covered = ("bbox_attention.", "mask_head.")
        # parameters that no stage owns (query_embed: its gradient comes out of plain torch autograd) are exchanged
        # one by one from a post-accumulate hook, on the same side stream
        if rt.stages is None:
            rt.build(inner)
        staged = {id(p) for st in rt.stages.values() for p in st.params}
        self._loose = [(n, p) for n, p in module.named_parameters()
                       if id(p) not in staged and not n.startswith(covered) and p.requires_grad]
        for n, p in self._loose:
            p.register_post_accumulate_grad_hook(lambda q, n=n: self.grad_sync.reduce(n, q.grad, ()))
        if is_dist_avail_and_initialized() and get_world_size() > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=0, group=process_group)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    @contextlib.contextmanager
    def no_sync(self):
        prev, self.grad_sync.enabled = self.grad_sync.enabled, False
        try:
            yield
        finally:
            self.grad_sync.enabled = prev
