"""GPU phaser: drop-in for the DSP of ``PedalboardPhaserDataset.apply_pedalboard_phaser``
(reference mod_extraction/datasets.py:455-482), which calls ``pedalboard.Phaser`` (JUCE dsp::Phaser).

PARITY UNPINNED: ``pedalboard==0.7.3`` is a third-party C++ wheel that is not available offline
(SURVEY F1).  The kernels follow this repo's own restatement of the JUCE algorithm and are
checked against its CPU version only; results must not be reported as "matches the reference".
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _ops, util

__all__ = ["Phaser", "apply_pedalboard_phaser"]


class Phaser(nn.Module):
    """Batched phaser.  ``forward(x, rate_hz, depth, centre_frequency_hz, feedback, mix)`` with the
    keyword names of ``pedalboard.Phaser``; x is (B, N) or (B, 1, N); parameters are floats or (B,)."""

    def __init__(self, sr: float = 44100.0, buffer_size: int = 8192) -> None:
        super().__init__()
        self.sr = sr
        self.buffer_size = buffer_size      # pedalboard processes in host blocks of this many samples

    @torch.no_grad()
    def forward(self, x: Tensor, rate_hz, depth, centre_frequency_hz, feedback, mix,
                example_index: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
        squeeze = x.ndim == 3
        if squeeze:
            assert x.size(1) == 1, "mono only (the reference datasets are mono)"
        on_cpu = not x.is_cuda
        if on_cpu:
            if not torch.cuda.is_available():
                raise RuntimeError("mod_extraction_b200 needs a CUDA device (no CPU fallback)")
            x = x.cuda(non_blocking=True)
        x2 = x.detach().float().reshape(x.size(0), x.size(-1))
        o2 = None if out is None else out.reshape(x2.shape)
        y = _ops.phaser(x2, self.sr, rate_hz, depth, centre_frequency_hz, feedback, mix, self.buffer_size,
                        example_index, o2)
        y = y.reshape(x.shape)
        return y.cpu() if on_cpu else y


def apply_pedalboard_phaser(x: Tensor, sr: float, rate_hz: float,
                            ranges: Dict[str, Dict[str, float]]) -> Tuple[Tensor, Dict[str, float]]:
    """datasets.py:455-482: same host RNG draws in the same order (depth, centre frequency,
    feedback, mix), same returned ``fx_params`` dict, same final clip -- the DSP runs on the GPU."""
    depth = util.sample_uniform(ranges["depth"]["min"], ranges["depth"]["max"])
    centre_frequency_hz = util.sample_log_uniform(ranges["centre_frequency_hz"]["min"],
                                                  ranges["centre_frequency_hz"]["max"])
    feedback = util.sample_uniform(ranges["feedback"]["min"], ranges["feedback"]["max"])
    mix = util.sample_uniform(ranges["mix"]["min"], ranges["mix"]["max"])
    assert x.ndim == 2
    y = Phaser(sr)(x, rate_hz, depth, centre_frequency_hz, feedback, mix)
    fx_params = {
        "depth": depth,
        "feedback": feedback,
        "mix": mix,
        "rate_hz": rate_hz,
        "shape": "cos",
    }
    return y, fx_params
