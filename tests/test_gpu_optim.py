"""Multi-tensor optimizer kernels (csrc/optim.cu) against torch's own implementations on the same fp32 tensors:
torch.optim.AdamW with three parameter groups (main.py:351-392), torch.nn.utils.clip_grad_norm_ (engine.py:89-90) and
the reference's update_ema formula (util/optim.py:9-26).  Tolerance 2e-6 relative (fp32, different FMA contraction)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def make_params(seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    shapes = [(256, 256), (2048,), (64, 3, 7, 7), (1,), (513, 7), (50265, 8), (3, 3), (1024, 1024)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]


def test_fused_adamw_matches_torch_adamw():
    from toist_b200.util.optim import FusedAdamW

    pa, pb = make_params(0), make_params(0)

    def groups(ps):
        return [{"params": ps[:3]}, {"params": ps[3:6], "lr": 1e-5}, {"params": ps[6:], "lr": 5e-5, "weight_decay": 0.0}]

    oa = torch.optim.AdamW(groups(pa), lr=1e-4, weight_decay=1e-4)
    ob = FusedAdamW(groups(pb), lr=1e-4, weight_decay=1e-4)
    g = torch.Generator(device="cpu").manual_seed(1)
    for step in range(5):
        for x, y in zip(pa, pb):
            if step == 2 and x.numel() == 1:  # a parameter without a gradient in one step keeps its state
                x.grad = y.grad = None
                continue
            gr = torch.randn(x.shape, generator=g).to(DEV) * (10.0 ** (step - 2))
            x.grad, y.grad = gr.clone(), gr.clone()
        if step == 3:
            for o in (oa, ob):
                o.param_groups[0]["lr"] = 3e-5  # adjust_learning_rate edits the groups between steps
        oa.step()
        ob.step()
        for x, y in zip(pa, pb):
            assert rel_err(y, x) < 2e-6, (step, tuple(x.shape))
    for x, y in zip(pa, pb):
        if x.numel() > 1:
            assert rel_err(ob.state[y]["exp_avg"], oa.state[x]["exp_avg"]) < 2e-6
            assert rel_err(ob.state[y]["exp_avg_sq"], oa.state[x]["exp_avg_sq"]) < 2e-6
    sd = ob.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


@pytest.mark.parametrize("scale", [1e-3, 1.0, 50.0])
def test_clip_grad_norm_matches_torch(scale):
    from toist_b200.util.optim import clip_grad_norm_

    pa, pb = make_params(2), make_params(2)
    g = torch.Generator(device="cpu").manual_seed(3)
    for x, y in zip(pa, pb):
        gr = torch.randn(x.shape, generator=g).to(DEV) * scale
        x.grad, y.grad = gr.clone(), gr.clone()
    pa[3].grad = pb[3].grad = None
    na = torch.nn.utils.clip_grad_norm_(pa, 0.1)
    nb = clip_grad_norm_(pb, 0.1)
    assert abs(float(na) - float(nb)) <= 2e-6 * float(na)
    for x, y in zip(pa, pb):
        if x.grad is not None:
            assert rel_err(y.grad, x.grad) < 2e-6
    n1, n2 = clip_grad_norm_(pb, 1e9), clip_grad_norm_(pb, 1e9)  # nothing is clipped: same gradients twice
    assert float(n1) == float(n2)  # fixed summation order: bit-reproducible
    assert abs(float(n1) - min(float(nb), 0.1)) <= 1e-5 * max(float(n1), 1e-3)


def test_update_ema_matches_reference_formula():
    from toist_b200.util.optim import update_ema

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(300, 77), torch.nn.BatchNorm1d(77), torch.nn.Linear(77, 5)).to(DEV)
    ema = copy.deepcopy(model)
    ref = copy.deepcopy(model)
    for step in range(3):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.1)
            model[1].running_mean.add_(0.5)
            model[1].num_batches_tracked.add_(1)
        update_ema(model, ema, 0.9998)
        with torch.no_grad():
            msd = model.state_dict()
            for k, v in ref.state_dict().items():
                v.copy_(v * 0.9998 + (1.0 - 0.9998) * msd[k].detach())
        for (k, a), (_, b) in zip(ema.state_dict().items(), ref.state_dict().items()):
            if a.dtype.is_floating_point:
                assert rel_err(a, b) < 1e-6, k
            else:
                assert torch.equal(a, b), k


def test_adamw_many_rows_and_intact_state_on_overflow():
    """Six parameter groups (the distillation recipe, main.py:351-386) with parameters lagging behind their group fit
    one launch; when a step would need more hyper-parameter rows than a launch carries, it raises BEFORE any step
    counter moves."""
    from toist_b200.util import optim as O

    torch.manual_seed(0)
    ps = [torch.randn(257, device="cuda", requires_grad=True) for _ in range(12)]
    ref = [p.detach().clone().requires_grad_(True) for p in ps]
    groups = lambda xs: [{"params": xs[2 * i: 2 * i + 2], "lr": 1e-3 * (i + 1)} for i in range(6)]  # noqa: E731
    o1, o2 = O.FusedAdamW(groups(ps), lr=1e-3, weight_decay=1e-2), torch.optim.AdamW(groups(ref), lr=1e-3, weight_decay=1e-2)
    for step in range(4):
        for i, (p, r) in enumerate(zip(ps, ref)):
            if i % 2 == 1 and step % 2 == 1:  # every second parameter skips every second step: 12 distinct rows
                p.grad = r.grad = None
                continue
            g = torch.randn(257, device="cuda")
            p.grad, r.grad = g.clone(), g.clone()
        o1.step()
        o2.step()
    for p, r in zip(ps, ref):
        assert torch.allclose(p, r, rtol=2e-6, atol=1e-7)
    # overflow: 40 parameters in one group, each with its own step count
    qs = [torch.randn(8, device="cuda", requires_grad=True) for _ in range(O.MAX_ADAM_ROWS + 8)]
    o3 = O.FusedAdamW(qs, lr=1e-3)
    for i, q in enumerate(qs):  # parameter i has taken i steps
        o3.state[q]["step"] = i
        o3.state[q]["exp_avg"] = torch.zeros_like(q)
        o3.state[q]["exp_avg_sq"] = torch.zeros_like(q)
        q.grad = torch.ones_like(q)
    before = [q.detach().clone() for q in qs]
    with pytest.raises(RuntimeError, match="state left untouched"):
        o3.step()
    assert [int(o3.state[q]["step"]) for q in qs] == list(range(len(qs)))
    assert all(torch.equal(a, q) for a, q in zip(before, qs))
