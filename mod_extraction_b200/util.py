"""Drop-in for the reference's ``mod_extraction/util.py``.

``linear_interpolate_last_dim`` (util.py:15-29) runs on the GPU; the RNG helpers (util.py:32-62)
stay on the host on purpose: "identical seeds" in the reference means the torch *global CPU*
generator (and numpy's global RNG behind scipy.stats.loguniform), so the draws must come from
the very same generators in the very same order.
"""
from __future__ import annotations

from typing import Any, List, Union

import torch as tr
from torch import Tensor as T

from . import _ops


def linear_interpolate_last_dim(x: T, n: int, align_corners: bool = True) -> T:
    """util.py:15-29 -- F.interpolate(mode="linear") along the last dim, bit-exact with torch CPU."""
    n_dim = x.ndim
    assert 1 <= n_dim <= 3
    if x.size(-1) == n:
        return x
    on_cpu = not x.is_cuda
    if on_cpu:
        if not tr.cuda.is_available():
            raise RuntimeError("mod_extraction_b200 needs a CUDA device (no CPU fallback)")
        xd = x.detach().float().cuda(non_blocking=True)
    else:
        xd = x.detach().float()
    y = _ops.interp_linear(xd, n, align_corners)
    return y.cpu() if on_cpu else y


def choice(items: List[Any]) -> Any:
    """util.py:32-35."""
    assert len(items) > 0
    return items[randint(0, len(items))]


def randint(low: int, high: int, n: int = 1) -> Union[int, T]:
    """util.py:38-42 (torch global CPU generator)."""
    x = tr.randint(low=low, high=high, size=(n,))
    return x.item() if n == 1 else x


def sample_uniform(low: float, high: float, n: int = 1) -> Union[float, T]:
    """util.py:45-49 (torch global CPU generator, float32 arithmetic)."""
    x = (tr.rand(n) * (high - low)) + low
    return x.item() if n == 1 else x


def sample_log_uniform(low: float, high: float, n: int = 1) -> Union[float, T]:
    """util.py:52-62 (scipy.stats.loguniform on numpy's global RNG)."""
    if low == high:
        return low if n == 1 else tr.full(size=(n,), fill_value=low)
    from scipy.stats import loguniform
    x = loguniform.rvs(low, high, size=n)
    return float(x[0]) if n == 1 else tr.from_numpy(x)
