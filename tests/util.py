"""Shared helpers for the test-suite (golden digests, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_SAMPLES = 64


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def digest(t: torch.Tensor) -> np.ndarray:
    """Same digest as tests/golden/make_golden.py: 64 evenly spaced samples + sum + abs-sum."""
    t = t.detach().double().flatten().cpu()
    n = t.numel()
    idx = torch.linspace(0, n - 1, min(N_SAMPLES, n), dtype=torch.float64).long().clamp_(max=n - 1)
    out = torch.zeros(N_SAMPLES + 2, dtype=torch.float64)
    out[: idx.numel()] = t[idx]
    out[-2] = t.sum()
    out[-1] = t.abs().sum()
    return out.numpy()


def assert_digest_close(actual: torch.Tensor, gold: np.ndarray, rtol: float, name: str = "",
                        floor: float = 0.0):
    """Relative-to-scale comparison: |a-g| <= rtol * max(max|g samples|, mean|g|, floor).

    Theoretically-zero gradients (softmax-invariant biases) are pure rounding noise in the
    reference too; ``floor`` gives them an absolute scale."""
    a = digest(actual)
    n = actual.numel()
    k = min(N_SAMPLES, n)
    scale = max(np.abs(gold[:k]).max(), gold[-1] / max(n, 1), floor, 1e-30)
    err = np.abs(a[:k] - gold[:k]).max()
    assert err <= rtol * scale, f"{name}: sample err {err:.3e} > {rtol:g} * scale {scale:.3e}"
    # abs-sum is a robust whole-tensor check
    assert abs(a[-1] - gold[-1]) <= rtol * max(gold[-1], floor * n, 1e-30) * 4 + 1e-12, \
        f"{name}: abs-sum {a[-1]:.6e} vs {gold[-1]:.6e}"


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
