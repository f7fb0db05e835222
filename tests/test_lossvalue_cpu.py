"""Host logic of the fused loss sum (toist_b200/models/lossvalue.py): the caller's `sum(loss_dict[k] * weight_dict[k] ...)`
and `.backward()` (reference engine.py:72,88) on LossValue entries must give exactly what plain 0-dim tensors give -
value, gradient, requires_grad - and every other use of an entry must behave like the tensor it stands for.  Pure
host arithmetic: runs on CPU tensors (nothing here launches a kernel)."""
from __future__ import annotations

import torch

from toist_b200.models.lossvalue import LossValue, loss_cells

ROWS, L = 5, 6
GRAD_ROWS = (True, True, True, False, True)  # cardinality_error carries no gradient (models/mdetr.py:783)


def _setup(seed=0):
    g = torch.Generator().manual_seed(seed)
    leaf = torch.randn(ROWS, L, generator=g, requires_grad=True)
    out = leaf * 2.0 + 1.0  # a non-leaf with grad_fn, like the criterion's output
    _, cells = loss_cells(out, GRAD_ROWS)
    names = [f"t{r}_{l}" for r in range(ROWS) for l in range(L)]
    weights = {n: 0.5 + 0.25 * i for i, n in enumerate(names) if (i // L) != 3}  # no weight for the cardinality row
    return leaf, out, dict(zip(names, cells)), weights


def test_weighted_sum_and_backward_match_plain_tensors():
    leaf, out, loss_dict, wd = _setup()
    total = sum(loss_dict[k] * wd[k] for k in loss_dict.keys() if k in wd)  # engine.py:72 verbatim
    assert type(total) is LossValue and total.requires_grad
    plain = sum(out.reshape(-1)[i] * wd[k] for i, k in enumerate(loss_dict) if k in wd)
    assert abs(float(total.detach()) - float(plain.detach())) <= 1e-5 * abs(float(plain.detach()))
    total.backward(retain_graph=True)  # (the comparison below walks the same graph once more)
    got = leaf.grad.clone()
    leaf.grad = None
    plain.backward()
    assert torch.allclose(got, leaf.grad, rtol=1e-6, atol=1e-7)
    assert float(got[3].abs().sum()) == 0.0  # the unweighted row receives nothing


def test_scalar_on_either_side_zero_start_and_extra_tensor():
    leaf, out, loss_dict, wd = _setup(1)
    a, b = loss_dict["t0_0"], loss_dict["t1_2"]
    extra_leaf = torch.tensor(3.0, requires_grad=True)
    extra = extra_leaf * 4.0  # an ordinary 0-dim tensor in the same sum (loss_cluster_feature in the distillation recipe)
    total = 0 + 2.0 * a + b * 3 + extra
    assert type(total) is LossValue
    want = 2.0 * out[0, 0] + 3.0 * out[1, 2] + extra
    assert abs(float(total.detach()) - float(want.detach())) < 1e-5
    total.backward()
    assert abs(float(extra_leaf.grad) - 4.0) < 1e-6
    expect = torch.zeros(ROWS, L)
    expect[0, 0], expect[1, 2] = 2.0 * 2.0, 3.0 * 2.0  # d out / d leaf = 2
    assert torch.allclose(leaf.grad, expect)


def test_every_other_use_materialises_an_ordinary_tensor():
    leaf, out, loss_dict, wd = _setup(2)
    v = loss_dict["t2_3"]
    assert abs(v.item() - float(out[2, 3])) < 1e-6 and abs(float(v) - float(out[2, 3])) < 1e-6
    st = torch.stack([loss_dict["t0_0"], loss_dict["t4_5"]], dim=0)  # util/dist.reduce_dict stacks the entries
    assert type(st) is torch.Tensor and torch.allclose(st.detach(), torch.stack([out[0, 0], out[4, 5]]).detach())
    assert bool(torch.isfinite(v)) and (v == v).item()
    d = v.detach()
    assert not d.requires_grad and abs(float(d) - float(out[2, 3])) < 1e-6
    prod = loss_dict["t0_1"] * loss_dict["t1_1"]  # not a linear combination: falls back to torch
    assert type(prod) is torch.Tensor and abs(float(prod.detach()) - float((out[0, 1] * out[1, 1]).detach())) < 1e-5
    assert not loss_dict["t3_0"].requires_grad and loss_dict["t0_0"].requires_grad  # per-row gradient flags
    assert "LossValue" in type(v).__name__ and isinstance(repr(v), str)


def test_backward_with_arguments_and_without_graph():
    leaf, out, loss_dict, wd = _setup(3)
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    total.backward(retain_graph=True)
    g1 = leaf.grad.clone()
    leaf.grad = None
    total.backward(gradient=torch.tensor(2.0))  # explicit gradient: the materialised path
    assert torch.allclose(leaf.grad, 2.0 * g1, rtol=1e-6)
    with torch.no_grad():
        _, cells = loss_cells((leaf * 2.0).detach(), GRAD_ROWS)
    s = cells[0] * 1.0 + cells[1] * 2.0
    assert not s.requires_grad
    try:
        s.backward()
    except RuntimeError as e:
        assert "does not require grad" in str(e)
    else:
        raise AssertionError("backward() of a value without graph must raise like torch does")
