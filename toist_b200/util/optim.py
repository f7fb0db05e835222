"""Optimizer side of the training step on the multi-tensor kernels of csrc/optim.cu.

Drop-ins, same call signatures as what the reference uses:
  * `FusedAdamW(param_dicts, lr=..., weight_decay=...)`   for `torch.optim.AdamW` (main.py:351-392), one launch per step
  * `clip_grad_norm_(parameters, max_norm)`               for `torch.nn.utils.clip_grad_norm_` (engine.py:89-90), three launches
  * `update_ema(model, model_ema, decay)`                 for util/optim.py:9-26, one launch
  * `adjust_learning_rate(optimizer, epoch, curr_step, num_training_steps, args)`  util/optim.py:29-90 (host arithmetic)
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from bisect import bisect_right
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from .. import _lib
from .misc import h2d

_CHUNK = 2048
MAX_ADAM_ROWS = 32  # csrc/optim.cu kMaxAdamRows: (parameter group, step count) rows one launch can carry


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Table:
    """Device array of OptItem records for a fixed list of tensor tuples; rebuilt only when a pointer changes (the
    gradient arenas of the CUDA-graph backward are static, parameters and optimizer state never move)."""

    def __init__(self):
        self.key = None
        self.dev = None
        self.blocks = 0
        self.n = 0

    def get(self, cols: Sequence[Sequence[Optional[torch.Tensor]]], groups: Optional[Sequence[int]] = None):
        """cols: up to four parallel lists (a, b, c, d) of contiguous fp32 CUDA tensors of equal numel per row."""
        ptrs = [[0 if t is None else t.data_ptr() for t in col] for col in cols]
        numel = [t.numel() for t in cols[0]]
        key = (tuple(map(tuple, ptrs)), tuple(numel), None if groups is None else tuple(groups))
        if key != self.key:
            n = len(numel)
            rec = np.zeros((n, 6), dtype=np.int64)
            for j, col in enumerate(ptrs):
                rec[:, j] = col
            rec[:, 4] = numel
            blocks = 0
            for i, ne in enumerate(numel):
                g = 0 if groups is None else int(groups[i])
                rec[i, 5] = blocks | (g << 32)  # {int32 first_block, int32 group}, little endian
                blocks += -(-ne // _CHUNK)
            for col in cols:
                for t in col:
                    if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                        raise RuntimeError("toist_b200.util.optim works on contiguous fp32 CUDA tensors (no CPU path)")
            assert _lib.load().toist_sizeof_opt_item() == 48
            self.dev = h2d(torch.from_numpy(rec.view(np.uint8).reshape(-1)), cols[0][0].device)
            self.key, self.blocks, self.n = key, blocks, n
        return self.dev, self.n, self.blocks


_clip_tables = {}


def clip_grad_norm_(parameters, max_norm: float, norm_type: float = 2.0) -> torch.Tensor:
    """Clips the global L2 norm of the gradients in place; returns the total norm (a 0-dim CUDA tensor, no host sync)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    if float(norm_type) != 2.0:
        raise NotImplementedError("only the L2 norm the reference uses (engine.py:90) is implemented")
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros(())
    dev = grads[0].device
    key = (dev.type, dev.index, len(grads))
    tab = _clip_tables.get(key)
    if tab is None:
        if len(_clip_tables) >= 16:  # a handful of (device, parameter count) combinations at most: stay bounded
            _clip_tables.clear()
        tab = _clip_tables[key] = (_Table(), {})
    items, n, blocks = tab[0].get([grads])
    if len(tab[1]) > 4:
        tab[1].clear()
    scratch = tab[1].get(blocks)
    if scratch is None:
        scratch = tab[1][blocks] = (torch.empty(blocks, dtype=torch.float32, device=dev),)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    L = _lib.load()
    _lib.check(L.toist_grad_sqnorm(items.data_ptr(), n, blocks, scratch[0].data_ptr(), float(max_norm), out.data_ptr(),
                                   _stream()))
    _lib.check(L.toist_grad_clip_scale(items.data_ptr(), n, blocks, out.data_ptr(), _stream()))
    return out[0]


_ema_tables = {}


def update_ema(model, model_ema, decay: float) -> None:
    """w_ema = w_ema * decay + (1 - decay) * w for every floating-point entry of the state dict (util/optim.py:9-26);
    integer buffers are copied."""
    with torch.no_grad():
        if hasattr(model, "module"):
            model = model.module
        key = (id(model), id(model_ema))
        ent = _ema_tables.get(key)
        probe = (next(model.parameters()).data_ptr(), next(model_ema.parameters()).data_ptr())
        if ent is not None and (ent[3]() is not model or ent[4]() is not model_ema or ent[5] != probe):
            ent = None  # the ids were recycled by other objects, or the parameters moved (.to(), .cuda())
        if ent is None:
            msd = model.state_dict()
            pairs, copies = [], []
            for k, ema_v in model_ema.state_dict().items():
                mv = msd[k].detach()
                if ema_v.dtype == torch.float32 and mv.dtype == torch.float32 and ema_v.is_contiguous() and mv.is_contiguous():
                    pairs.append((ema_v, mv))
                else:
                    copies.append((ema_v, mv))
            ent = _ema_tables[key] = (_Table(), pairs, copies, weakref.ref(model), weakref.ref(model_ema), probe)
        tab, pairs, copies = ent[:3]
        if pairs:
            items, n, blocks = tab.get([[a for a, _ in pairs], [b for _, b in pairs]])
            _lib.check(_lib.load().toist_ema_update(items.data_ptr(), n, blocks, float(decay), 1.0 - float(decay), _stream()))
        for ema_v, mv in copies:
            ema_v.copy_(ema_v * decay + (1.0 - decay) * mv)


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, no amsgrad) with the whole step of all
    parameter groups in ONE launch.  State (`step`, `exp_avg`, `exp_avg_sq`) uses torch's key names."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > MAX_ADAM_ROWS:
            raise ValueError(f"FusedAdamW supports up to {MAX_ADAM_ROWS} parameter groups")
        self._table = _Table()

    def zero_grad(self, set_to_none: bool = True) -> None:
        """engine.py:86.  With `set_to_none` (torch's default) this is one attribute store per parameter; torch's own
        loop (per-parameter hook / foreach bookkeeping) costs 1.5 ms per step for this model's ~950 tensors."""
        if not set_to_none:
            return super().zero_grad(set_to_none=False)
        for group in self.param_groups:
            for p in group["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ps, gs, ms, vs, grp = [], [], [], [], []
        rows = {}  # (group index, step count) -> hyper-parameter row: parameters skipped in some steps lag behind
        hyper = []
        bump = []  # step counters move only after the row count has been validated (state stays intact on error)
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                t = int(st["step"]) + 1
                bump.append(st)
                row = rows.get((gi, t))
                if row is None:
                    row = rows[(gi, t)] = len(hyper)
                    bc1, bc2 = 1.0 - b1 ** t, 1.0 - b2 ** t
                    hyper.append((1.0 - group["lr"] * group["weight_decay"], 1.0 - b1, b2, 1.0 - b2, group["eps"],
                                  group["lr"] / bc1, math.sqrt(bc2), 0.0))
                ps.append(p)
                gs.append(p.grad)
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
                grp.append(row)
        if len(hyper) > MAX_ADAM_ROWS:
            raise RuntimeError(f"FusedAdamW: {len(hyper)} distinct (parameter group, step count) combinations in one "
                               f"step (limit {MAX_ADAM_ROWS}); optimizer state left untouched")
        for st in bump:
            st["step"] += 1
        hyper = np.asarray(hyper, dtype=np.float32).reshape(-1, 8)
        if ps:
            items, n, blocks = self._table.get([ps, gs, ms, vs], grp)
            _lib.check(_lib.load().toist_adamw_step(items.data_ptr(), n, blocks, hyper.ctypes.data_as(C.c_void_p),
                                                    int(hyper.shape[0]), _stream()))
            # the kernel writes through raw pointers: tell autograd (and the models' bf16 shadow-weight cache, which
            # refreshes when a master's version moved: runtime.ShadowBank.ensure) that the parameters changed
            torch.autograd.graph.increment_version(ps)
        return loss


def _schedule_gammas(epoch: int, curr_step: int, num_training_steps: int, args):
    """(gamma of transformer / backbone, gamma of the text encoder) for the four schedules of util/optim.py:29-90."""
    warm = round(args.fraction_warmup_steps * num_training_steps)

    def linear():
        if curr_step < warm:
            return float(curr_step) / float(max(1, warm))
        return max(0.0, float(num_training_steps - curr_step) / float(max(1, num_training_steps - warm)))

    if args.schedule == "step":
        gamma = text_gamma = 0.1 ** (epoch // args.lr_drop)
    elif args.schedule == "multistep":
        gamma = text_gamma = 0.5 ** bisect_right(list(range(args.lr_drop, args.epochs, 50)), epoch)
    elif args.schedule == "linear_with_warmup":
        gamma, text_gamma = 0.1 ** (epoch // args.lr_drop), linear()
    elif args.schedule == "all_linear_with_warmup":
        gamma = text_gamma = linear()
    else:
        raise NotImplementedError(args.schedule)
    return gamma, text_gamma


def adjust_learning_rate(optimizer, epoch: int, curr_step: int, num_training_steps: int, args) -> None:
    """util/optim.py:29-90: three parameter groups (transformer + heads, backbone, text encoder; main.py:351-367)."""
    gamma, text_gamma = _schedule_gammas(epoch, curr_step, num_training_steps, args)
    base = [args.lr, args.lr_backbone, args.text_encoder_lr]
    assert len(optimizer.param_groups) == len(base)
    for group, lr, g in zip(optimizer.param_groups, base, [gamma, gamma, text_gamma]):
        group["lr"] = lr * g


def dis_adjust_learning_rate(optimizer, epoch: int, curr_step: int, num_training_steps: int, args) -> None:
    """util/optim.py:92-152: the distillation recipe optimises the student and the noun (teacher) model together, six
    parameter groups = the three groups of `adjust_learning_rate` once per model (main.py:368-386)."""
    gamma, text_gamma = _schedule_gammas(epoch, curr_step, num_training_steps, args)
    base = [args.lr, args.lr_backbone, args.text_encoder_lr] * 2
    assert len(optimizer.param_groups) == len(base)
    for group, lr, g in zip(optimizer.param_groups, base, [gamma, gamma, text_gamma] * 2):
        group["lr"] = lr * g
