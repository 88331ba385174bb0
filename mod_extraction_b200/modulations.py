"""Drop-in for the LFO generators of the reference's ``mod_extraction/modulations.py``.

All float arithmetic runs in libmodfx.so on the GPU; the host keeps only what must stay there:
argument checks, the torch-global-RNG draws in the reference's order and integer
bookkeeping of section boundaries.  Signatures and assertion behaviour follow the reference.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple, Union

import torch as tr
from torch import Tensor as T

from . import _ops, util
from ._lib import SHAPE_ID, SHAPES

__all__ = ["SHAPES", "SHAPE_ID", "make_mod_signal", "make_mod_signal_batch", "make_rand_mod_signal",
           "lfo_kernel_params"]


def _device() -> tr.device:
    if not tr.cuda.is_available():
        raise RuntimeError("mod_extraction_b200 needs a CUDA device (no CPU fallback)")
    return tr.device("cuda", tr.cuda.current_device())


def shape_ids(shapes: Union[T, Sequence[str], Sequence[int]]) -> T:
    if isinstance(shapes, T):
        return shapes.to(tr.int32)
    return tr.tensor([SHAPE_ID[s] if isinstance(s, str) else int(s) for s in shapes], dtype=tr.int32)


def lfo_kernel_params(freq, phase, shape, exp, sr: float, device) -> Tuple[T, T, T, Optional[T]]:
    """Host-side preparation of make_mod_signal's arguments for the kernels (modulations.py:22-30):
    the reference's asserts, and the halving of freq / phase for the rectified shapes."""
    f = tr.as_tensor(freq, dtype=tr.float64).reshape(-1).cpu()
    p = tr.as_tensor(phase, dtype=tr.float64).reshape(-1).cpu()
    s = shape_ids(shape).reshape(-1).cpu()
    assert bool((f > 0.0).all()) and bool((f < sr / 2.0).all())            # modulations.py:23
    assert bool((p >= -2 * tr.pi).all()) and bool((p <= 2 * tr.pi).all())  # modulations.py:24
    assert bool((s >= 0).all()) and bool((s < len(SHAPES)).all())          # modulations.py:25
    rect = (s == SHAPE_ID["rect_cos"]) | (s == SHAPE_ID["inv_rect_cos"])
    f = tr.where(rect, f / 2.0, f)                                         # modulations.py:26-29
    p = tr.where(rect, p / 2.0, p)
    e = None
    if exp is not None:
        e = tr.as_tensor(exp, dtype=tr.float64).reshape(-1).cpu()
        if e.numel() == 1 and f.numel() > 1:
            e = e.expand(f.numel())
        assert bool((e > 0).all())                                         # modulations.py:30
        e = e.float().to(device, non_blocking=True)
    return (f.float().to(device, non_blocking=True), p.float().to(device, non_blocking=True),
            s.to(device, non_blocking=True), e)


def make_mod_signal_batch(n_samples: int, sr: float, freq, phase, shape, exp=None, device=None) -> T:
    """Batched make_mod_signal: one launch for B LFOs.  Returns (B, n_samples) on the GPU."""
    assert n_samples > 0
    device = _device() if device is None else device
    f, p, s, e = lfo_kernel_params(freq, phase, shape, exp, sr, device)
    return _ops.lfo(n_samples, sr, f, p, s, e)


def make_mod_signal(n_samples: int,
                    sr: float,
                    freq: float,
                    phase: float = 0.0,
                    shape: str = "cos",
                    exp: float = 1.0,
                    device=None) -> T:
    """modulations.py:16-57.  Returns a (n_samples,) tensor on the GPU."""
    assert n_samples > 0
    assert 0.0 < freq < sr / 2.0
    assert -2 * tr.pi <= phase <= 2 * tr.pi
    assert shape in {"cos", "rect_cos", "inv_rect_cos", "tri", "saw", "rsaw", "sqr"}
    assert exp > 0
    return make_mod_signal_batch(n_samples, sr, [float(freq)], [float(phase)], [shape],
                                 None if exp == 1.0 else [float(exp)], device)[0]


def make_rand_mod_signal(batch_size: int,
                         n_samples: int,
                         sr: float,
                         freq_min: float,
                         freq_max: float,
                         shapes_gt: Optional[Sequence] = None,
                         shapes: Optional[List[str]] = None,
                         phase_gt: Optional[T] = None,
                         phase_error: float = 0.5,
                         freq_gt: Optional[T] = None,
                         freq_error: float = 0.25,
                         device=None) -> T:
    """modulations.py:60-101: same host RNG draws in the same order (phase, freq, shape per
    example), then ONE kernel launch for the whole batch instead of a python loop of LFOs."""
    if shapes is None:
        shapes = ["cos", "tri", "rect_cos", "inv_rect_cos", "saw", "rsaw"]
    phases, freqs, shape_list = [], [], []
    for idx in range(batch_size):
        if phase_gt is not None:
            assert phase_gt.size(0) == batch_size
            phase = float(phase_gt[idx])
            if phase_error > 0:
                error = util.sample_uniform(-1.0, 1.0) * tr.pi * phase_error
                phase += error
                phase = (phase + (2 * tr.pi)) % (2 * tr.pi)
        else:
            phase = util.sample_uniform(0.0, 2 * tr.pi)
        if freq_gt is not None:
            assert freq_gt.size(0) == batch_size
            freq = float(freq_gt[idx])
            if freq_error > 0:
                error = util.sample_uniform(1.0 - freq_error, 1.0 + freq_error)
                freq *= error
                freq = min(max(freq, freq_min), freq_max)
        else:
            freq = util.sample_uniform(freq_min, freq_max)
        if shapes_gt is not None:
            assert len(shapes_gt) == batch_size
            shape = shapes_gt[idx]
        else:
            shape = util.choice(shapes)
        phases.append(phase)
        freqs.append(freq)
        shape_list.append(shape)
    return make_mod_signal_batch(n_samples, sr, freqs, phases, shape_list, None, device)
