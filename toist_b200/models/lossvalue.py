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
    __slots__ = ("out", "flat", "plain")

    def __init__(self, out: torch.Tensor):
        self.out = out
        self.flat = out.reshape(-1)
        self.plain = self.flat.detach()  # values without autograd history: storage behind the LossValue objects


_NUM = (int, float)
_make = torch.Tensor._make_subclass


class LossValue(torch.Tensor):
    """_lv = (terms {(id(bundle), cell index): coefficient}, bundles {id: bundle}, extras [(tensor, coef)], requires_grad).
    The tensor storage behind the object is a placeholder (the first cell it refers to): every read goes through
    __torch_function__, which materialises the real value."""

    @staticmethod
    def cell(bundle: _Bundle, index: int, requires_grad: bool) -> "LossValue":
        plain = bundle.plain[index]
        v = _make(LossValue, plain)
        v._lv = ({(id(bundle), index): 1.0}, {id(bundle): bundle}, (), requires_grad)
        v._plain = plain
        return v

    def _derive(self, terms, bundles, extras, rg) -> "LossValue":
        v = _make(LossValue, self._plain)  # shares the placeholder storage: no kernel, no copy
        v._lv = (terms, bundles, extras, rg)
        v._plain = self._plain
        return v

    # ------------------------------------------------------------------ symbolic algebra
    def _scaled(self, a: float) -> "LossValue":
        terms, bundles, extras, rg = self._lv
        return self._derive({k: c * a for k, c in terms.items()}, bundles, tuple((t, c * a) for t, c in extras), rg)

    def _plus(self, other) -> "LossValue":
        terms, bundles, extras, rg = self._lv
        if isinstance(other, LossValue):
            ot, ob, oe, org = other._lv
            t2 = dict(terms)
            for k, c in ot.items():
                t2[k] = t2.get(k, 0.0) + c
            b2 = bundles if ob.keys() <= bundles.keys() else {**bundles, **ob}
            return self._derive(t2, b2, extras + oe, rg or org)
        # an ordinary tensor (e.g. loss_cluster_feature): carried along, added when materialised / differentiated
        return self._derive(terms, bundles, extras + ((other, 1.0),), rg or bool(other.requires_grad))

    def _coef_vectors(self):
        """[(bundle, host fp32 coefficient vector over its cells)]"""
        terms, bundles, _, _ = self._lv
        vecs = {}
        for (bid, i), c in terms.items():
            w = vecs.get(bid)
            if w is None:
                w = vecs[bid] = torch.zeros(bundles[bid].flat.numel(), dtype=torch.float32)
            w[i] = c
        return [(bundles[bid], w) for bid, w in vecs.items()]

    def materialize(self) -> torch.Tensor:
        """The ordinary (differentiable) tensor this value stands for."""
        terms, bundles, extras, rg = self._lv
        total = None
        if len(terms) == 1 and not extras:
            ((bid, i), c), = terms.items()
            total = bundles[bid].flat[i]
            if c != 1.0:
                total = total * c
        else:
            for b, w in self._coef_vectors():
                part = (b.flat * h2d(w, b.flat.device)).sum()
                total = part if total is None else total + part
            for t, c in extras:
                part = t if c == 1.0 else t * c
                total = part if total is None else total + part
        return total if rg else total.detach()

    def _backward(self, gradient=None, retain_graph=None, create_graph=False, inputs=None):
        _, _, extras, _ = self._lv
        if gradient is not None or create_graph or inputs is not None:
            return self.materialize().backward(gradient, retain_graph, create_graph, inputs=inputs)
        roots, grads = [], []
        for b, w in self._coef_vectors():
            if b.out.requires_grad:
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
        name = getattr(func, "__name__", "")
        if not kwargs and len(args) == 2:
            a, b = args
            if name in ("mul", "__mul__", "__rmul__"):
                if type(a) is LossValue and isinstance(b, _NUM):
                    return a._scaled(float(b))
                if type(b) is LossValue and isinstance(a, _NUM):
                    return b._scaled(float(a))
            elif name in ("add", "__add__", "__radd__"):
                if type(b) is LossValue and type(a) is not LossValue:
                    a, b = b, a
                if type(a) is LossValue:
                    if isinstance(b, _NUM):
                        if b == 0:
                            return a  # sum() starts from 0
                    elif isinstance(b, torch.Tensor) and b.dim() == 0:
                        return a._plus(b)
        if name == "backward" and type(args[0]) is LossValue:
            return args[0]._backward(*args[1:], **(kwargs or {}))
        if name == "__get__" and len(args) == 1 and type(args[0]) is LossValue \
                and getattr(func, "__self__", None) is torch.Tensor.requires_grad:
            return args[0]._lv[3]
        kwargs = kwargs or {}
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
