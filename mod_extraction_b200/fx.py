"""Drop-in for the reference's ``mod_extraction/fx.py`` backed by the sm_100a kernels.

Same names, call signatures and argument meaning as the reference (fx.py:13-22 ``apply_tremolo``,
fx.py:25-130 ``MonoFlangerChorusModule``); shape / range violations raise ``AssertionError``
exactly where the reference asserts.  Differences, all deliberate:

* the work is done by ``libmodfx.so`` on the GPU.  CUDA inputs give a CUDA result; CPU inputs
  (the situation at the reference's call site ``FlangerCPUDataModule.on_before_batch_transfer``,
  data_modules.py:457) are staged through pinned memory and the result is returned on the CPU.
  There is no CPU implementation: without the library or a GPU the call raises ``RuntimeError``.
* no persistent ``delay_buf`` / ``out_buf`` (fx.py:43-44): the delay line lives in shared memory
  and is zeroed per call like fx.py:92-93, so the module is re-entrant and stream-safe.
* ``forward_control_rate`` / ``forward_lfo`` are additions: they fuse the x100 LFO upsample
  (data_modules.py:454-455) and the LFO synthesis (datasets.py:382) into the effect kernel so the
  audio-rate modulation signal never exists in HBM.
"""
from __future__ import annotations

from typing import Optional, Union

import torch
from torch import Tensor, nn

from . import _ops
from ._ops import ModSource

Param = Union[float, Tensor]


def _delay_samples(ms: float, sr: float) -> int:
    return int(((ms / 1000.0) * sr) + 0.5)      # fx.py:40-41


def _to_cuda(t: Tensor, device: torch.device) -> Tensor:
    if t.is_cuda:
        return t
    if not t.is_pinned():
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("mod_extraction_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _check_param(param: Param, bs: int, can_be_one: bool = True) -> None:
    """fx.py:46-70 (the view() reshapes are not needed here)."""
    if isinstance(param, Tensor):
        assert param.shape == (bs,)
        assert param.min() >= 0
        if can_be_one:
            assert param.max() <= 1.0
        else:
            assert param.max() < 1.0
    else:
        assert param >= 0
        if can_be_one:
            assert param <= 1.0
        else:
            assert param < 1.0


def apply_tremolo(x: Tensor, mod_sig: Tensor, mix: Param = 1.0) -> Tensor:
    """fx.py:13-22."""
    assert x.ndim == 3
    assert x.size(0) == mod_sig.size(0)
    assert x.size(-1) == mod_sig.size(-1)
    if isinstance(mix, Tensor):
        assert mix.size(0) == x.size(0)
        assert bool((0.0 <= mix).all()) and bool((mix <= 1.0).all())
    else:
        assert 0.0 <= mix <= 1.0
    on_cpu = not x.is_cuda
    dev = _default_device() if on_cpu else x.device
    xd = _to_cuda(x.detach().float(), dev)
    md = _to_cuda(mod_sig.detach().float(), dev)
    y = _ops.tremolo(xd, ModSource.audio_rate(md), mix)
    return y.cpu() if on_cpu else y


class MonoFlangerChorusModule(nn.Module):
    """Reference: fx.py:25-130.  Flanger and chorus differ only in ``max_min_delay_ms``."""

    def __init__(self,
                 batch_size: int,
                 n_ch: int,
                 n_samples: int,
                 sr: float,
                 max_min_delay_ms: float,
                 max_lfo_delay_ms: float,
                 check_ranges: bool = True,
                 interpolation: str = "linear") -> None:
        super().__init__()
        # "linear" is the reference (fx.py:113); "allpass" is an addition with no reference counterpart:
        # first-order all-pass fractional-delay interpolation, defined in include/modfx.h
        assert interpolation in ("linear", "allpass")
        self.interpolation = interpolation
        self.batch_size = batch_size
        self.n_ch = n_ch
        self.n_samples = n_samples
        self.sr = sr
        self.max_min_delay_ms = max_min_delay_ms
        self.max_lfo_delay_ms = max_lfo_delay_ms
        self.max_min_delay_samples = _delay_samples(max_min_delay_ms, sr)
        self.max_lfo_delay_samples = _delay_samples(max_lfo_delay_ms, sr)
        self.max_delay_samples = self.max_min_delay_samples + self.max_lfo_delay_samples
        # Range asserts on tensors cost a reduction (+ a sync for CUDA tensors), like fx.py:53-57.
        self.check_ranges = check_ranges

    # ------------------------------------------------------------------ internals
    def _checks(self, x: Tensor, feedback, min_delay_width, width, depth, mix) -> None:
        assert x.ndim == 3                                           # fx.py:80
        batch_size = x.size(0)
        assert x.shape == (self.batch_size, self.n_ch, self.n_samples), \
            "x must match the (batch_size, n_ch, n_samples) the module was built for (fx.py:43-44)"
        if self.check_ranges:
            _check_param(feedback, batch_size, can_be_one=False)     # fx.py:86-90
            _check_param(min_delay_width, batch_size)
            _check_param(width, batch_size)
            _check_param(depth, batch_size)
            _check_param(mix, batch_size)
        else:
            for p in (feedback, min_delay_width, width, depth, mix):
                if isinstance(p, Tensor):
                    assert p.shape == (batch_size,)

    def _render(self, x: Tensor, mod: ModSource, feedback, min_delay_width, width, depth, mix,
                example_index: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
        return _ops.flanger_chorus(x, mod, self.max_min_delay_samples, self.max_lfo_delay_samples,
                                   feedback, min_delay_width, width, depth, mix, example_index, out, self.interpolation)

    # ------------------------------------------------------------------ reference API
    def apply_effect(self,
                     x: Tensor,
                     mod_sig: Tensor,
                     feedback: Param,
                     min_delay_width: Param,
                     width: Param,
                     depth: Param,
                     mix: Param) -> Tensor:
        """fx.py:72-119."""
        assert x.ndim == 3
        batch_size, n_ch, n_samples = x.shape
        assert mod_sig.size(0) == batch_size                         # fx.py:82
        assert mod_sig.size(-1) == n_samples                         # fx.py:83
        self._checks(x, feedback, min_delay_width, width, depth, mix)
        on_cpu = not x.is_cuda
        dev = _default_device() if on_cpu else x.device
        xd = _to_cuda(x.detach().float(), dev)
        md = _to_cuda(mod_sig.detach().float(), dev)
        y = self._render(xd, ModSource.audio_rate(md), feedback, min_delay_width, width, depth, mix)
        return y.cpu() if on_cpu else y

    def forward(self,
                x: Tensor,
                mod_sig: Tensor,
                feedback: Param = 0.0,
                min_delay_width: Param = 1.0,
                width: Param = 1.0,
                depth: Param = 1.0,
                mix: Param = 1.0) -> Tensor:
        """fx.py:121-130."""
        with torch.no_grad():
            return self.apply_effect(x, mod_sig, feedback, min_delay_width, width, depth, mix)

    # ------------------------------------------------------------------ fused additions
    @torch.no_grad()
    def forward_control_rate(self, x: Tensor, mod_lo: Tensor, feedback: Param = 0.0, min_delay_width: Param = 1.0,
                             width: Param = 1.0, depth: Param = 1.0, mix: Param = 1.0,
                             example_index: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
        """``forward(x, linear_interpolate_last_dim(mod_lo, n_samples), ...)`` with the upsample
        (data_modules.py:454-455, util.py:15-29) fused into the kernel.  mod_lo: (B, n_lo)."""
        assert mod_lo.ndim == 2 and mod_lo.size(0) == x.size(0)
        self._checks(x, feedback, min_delay_width, width, depth, mix)
        on_cpu = not x.is_cuda
        dev = _default_device() if on_cpu else x.device
        xd = _to_cuda(x.detach().float(), dev)
        md = _to_cuda(mod_lo.detach().float(), dev)
        y = self._render(xd, ModSource.control_rate(md), feedback, min_delay_width, width, depth, mix,
                         example_index, out)
        return y.cpu() if on_cpu else y

    @torch.no_grad()
    def forward_lfo(self, x: Tensor, rate_hz: Tensor, phase: Tensor, shape: Tensor, exp: Optional[Tensor] = None,
                    n_lo: Optional[int] = None, sr_lo: Optional[float] = None, feedback: Param = 0.0,
                    min_delay_width: Param = 1.0, width: Param = 1.0, depth: Param = 1.0, mix: Param = 1.0,
                    example_index: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
        """Effect with the LFO synthesised in-kernel: equivalent to
        ``mod = stack([make_mod_signal(n_lo, sr_lo, rate_hz[b], phase[b], shape[b], exp[b])])``
        (datasets.py:382, defaults n_lo = n_samples // 100, sr_lo = sr // 100) followed by
        ``forward_control_rate``.  ``shape`` holds modfx shape ids (see modulations.SHAPE_ID)."""
        from .modulations import lfo_kernel_params
        n_lo = self.n_samples // 100 if n_lo is None else n_lo
        sr_lo = self.sr // 100 if sr_lo is None else sr_lo
        self._checks(x, feedback, min_delay_width, width, depth, mix)
        on_cpu = not x.is_cuda
        dev = _default_device() if on_cpu else x.device
        xd = _to_cuda(x.detach().float(), dev)
        f, p, s, e = lfo_kernel_params(rate_hz, phase, shape, exp, sr_lo, dev)
        y = self._render(xd, ModSource.lfo(n_lo, float(sr_lo), f, p, s, e), feedback, min_delay_width, width,
                         depth, mix, example_index, out)
        return y.cpu() if on_cpu else y
