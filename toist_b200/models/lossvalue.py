"""Loss-dictionary entries that stay symbolic through the caller's weighted sum (engine.py:72).

`SetCriterion.forward` computes every term of every decoder layer in a handful of launches into ONE tensor `out`
[5, L].  The reference's loop then does `sum(loss_dict[k] * weight_dict[k] for k in ...)` and `losses.backward()`:
with ordinary 0-dim tensors that is 24 multiplications + 24 additions issued one tiny kernel at a time, and as many
autograd nodes on the way back (~75 launches, 0.3 ms during which the GPU idles between the criterion and the backward
pass).  `LossValue` is a torch.Tensor subclass whose `* python-number` and `+ LossValue` stay on the host as a linear
combination {cell of `out`: coefficient}; `backward()` uploads the coefficients once and calls autograd on `out`
directly.  Every OTHER operation (`.item()`, `torch.stack` in reduce_dict, `float()`, comparisons, printing, ...)
first materialises an ordinary differentiable tensor and proceeds as torch would, so unknown callers keep working.

Opt-in: `SetCriterion.enable_fused_loss_sum()` (bench.py and the tools switch it on; the default hands out plain tensors).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from ..util.misc import h2d


class _Bundle:
    """One criterion output tensor [rows, L] (fp32, with grad_fn when anything upstream trains)."""
    __slots__ = ("out", "flat")

    def __init__(self, out: torch.Tensor):
        self.out = out
        self.flat = out.reshape(-1)


_NUM = (int, float)


class LossValue(torch.Tensor):
    @staticmethod
    def cell(bundle: _Bundle, index: int, requires_grad: bool) -> "LossValue":
        v = torch.Tensor._make_subclass(LossValue, bundle.flat.detach()[index])
        v._lv = ({id(bundle): (bundle, {index: 1.0})}, [], requires_grad)
        return v

    @staticmethod
    def _new(terms, extras, requires_grad, like: torch.Tensor) -> "LossValue":
        v = torch.Tensor._make_subclass(LossValue, like.detach().as_subclass(torch.Tensor).reshape(()))
        v._lv = (terms, extras, requires_grad)
        return v

    # ------------------------------------------------------------------ symbolic algebra
    def _scaled(self, a: float) -> "LossValue":
        terms, extras, rg = self._lv
        t2 = {k: (b, {i: c * a for i, c in d.items()}) for k, (b, d) in terms.items()}
        return LossValue._new(t2, [(t, c * a) for t, c in extras], rg, self)

    def _plus(self, other) -> "LossValue":
        terms, extras, rg = self._lv
        t2 = {k: (b, dict(d)) for k, (b, d) in terms.items()}
        e2 = list(extras)
        if isinstance(other, LossValue):
            ot, oe, org = other._lv
            for k, (b, d) in ot.items():
                if k in t2:
                    dst = t2[k][1]
                    for i, c in d.items():
                        dst[i] = dst.get(i, 0.0) + c
                else:
                    t2[k] = (b, dict(d))
            e2 += oe
            rg = rg or org
        else:  # an ordinary tensor (e.g. loss_cluster_feature): carried along, added when materialised / differentiated
            e2.append((other, 1.0))
            rg = rg or bool(other.requires_grad)
        return LossValue._new(t2, e2, rg, self)

    def materialize(self) -> torch.Tensor:
        """The ordinary (differentiable) tensor this value stands for."""
        terms, extras, rg = self._lv
        total = None
        for b, d in terms.values():
            idx = sorted(d)
            if len(idx) == 1 and d[idx[0]] == 1.0:
                part = b.flat[idx[0]]
            else:
                w = torch.zeros(b.flat.numel(), dtype=torch.float32)
                for i in idx:
                    w[i] = d[i]
                part = (b.flat * h2d(w, b.flat.device)).sum()
            total = part if total is None else total + part
        for t, c in extras:
            part = t if c == 1.0 else t * c
            total = part if total is None else total + part
        return total if rg else total.detach()

    def _backward(self, gradient=None, retain_graph=None, create_graph=False, inputs=None):
        terms, extras, rg = self._lv
        if gradient is not None or create_graph or inputs is not None:
            return self.materialize().backward(gradient, retain_graph, create_graph, inputs=inputs)
        roots, grads = [], []
        for b, d in terms.values():
            if not b.out.requires_grad:
                continue
            w = torch.zeros(b.flat.numel(), dtype=torch.float32)
            for i, c in d.items():
                w[i] = c
            roots.append(b.out)
            grads.append(h2d(w, b.out.device).view(b.out.shape))
        for t, c in extras:
            if t.requires_grad:
                roots.append(t)
                grads.append(torch.full_like(t, c))
        if not roots:
            raise RuntimeError("element 0 of tensors does not require grad and does not have a grad_fn")
        torch.autograd.backward(roots, grads, retain_graph=retain_graph)

    # ------------------------------------------------------------------ dispatch
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name in ("mul", "__mul__", "__rmul__") and len(args) == 2 and not kwargs:
            a, b = args
            if isinstance(a, LossValue) and isinstance(b, _NUM):
                return a._scaled(float(b))
            if isinstance(b, LossValue) and isinstance(a, _NUM):
                return b._scaled(float(a))
        elif name in ("add", "__add__", "__radd__") and len(args) == 2 and not kwargs:
            a, b = args
            if isinstance(b, LossValue) and not isinstance(a, LossValue):
                a, b = b, a
            if isinstance(a, LossValue):
                if isinstance(b, _NUM) and b == 0:
                    return a  # sum() starts from 0
                if isinstance(b, torch.Tensor) and b.dim() == 0:
                    return a._plus(b)
        elif name == "backward" and isinstance(args[0], LossValue):
            return args[0]._backward(*args[1:], **kwargs)
        elif name == "__get__" and len(args) == 1 and isinstance(args[0], LossValue):
            prop = getattr(func, "__self__", None)
            if prop is torch.Tensor.requires_grad:
                return args[0]._lv[2]
        with torch._C.DisableTorchFunctionSubclass():
            real = [a.materialize() if isinstance(a, LossValue) else
                    type(a)(x.materialize() if isinstance(x, LossValue) else x for x in a) if isinstance(a, (list, tuple)) else a
                    for a in args]
            rkw = {k: (v.materialize() if isinstance(v, LossValue) else v) for k, v in kwargs.items()}
            return func(*real, **rkw)


def loss_cells(out: torch.Tensor, requires_grad_rows) -> Tuple[_Bundle, List[LossValue]]:
    """All cells of `out` [rows, L] as LossValues; `requires_grad_rows[r]` says whether row r carries a gradient."""
    b = _Bundle(out)
    L = out.shape[1]
    return b, [LossValue.cell(b, i, bool(requires_grad_rows[i // L]) and out.requires_grad) for i in range(out.numel())]
