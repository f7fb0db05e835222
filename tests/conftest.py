"""pytest configuration: the `gpu` marker (tests that need a B200 and libtoist_b200.so) and shared helpers."""
from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a) and the built libtoist_b200.so")
    config.addinivalue_line("markers", "reference: needs /root/reference (this container only, never the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch

    have_gpu = torch.cuda.is_available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(skip_gpu)


def rel_err(got, ref) -> float:
    """Norm-wise relative error ||got - ref||_2 / ||ref||_2 in float64."""
    import torch

    g = got.detach().double().cpu().flatten()
    r = ref.detach().double().cpu().flatten()
    assert g.shape == r.shape, f"shape mismatch {tuple(got.shape)} vs {tuple(ref.shape)}"
    if not bool(torch.isfinite(g).all()):
        return float("inf")
    den = r.norm().item()
    return (g - r).norm().item() / (den if den > 0 else 1.0)


def max_err(got, ref) -> float:
    g = got.detach().double().cpu().flatten()
    r = ref.detach().double().cpu().flatten()
    return (g - r).abs().max().item() if g.numel() else 0.0


GRAD_SAMPLE = 2048


def strided_sample(t, n: int = GRAD_SAMPLE):
    """Deterministic subsample of a tensor (at most n elements, evenly strided over the flattened tensor): lets a
    small fixture pin a 170 M-element gradient set element by element (tools/make_golden.py, tests/test_gpu_golden.py)."""
    f = t.detach().flatten()
    if f.numel() <= n:
        return f.clone()
    step = f.numel() // n
    return f[::step][:n].clone()
