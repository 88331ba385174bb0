"""Registers the render ops with PyTorch as ``torch.ops.modfx.*`` -- CUDA dispatch key only.

There is deliberately no CPU (or any other backend) registration: calling an op with CPU tensors
fails in the dispatcher ("could not run 'modfx::...' with arguments from the 'CPU' backend"), which is the
"no CPU fallback" rule enforced at the framework level.  The implementations are the ctypes calls into
libmodfx.so of ``_ops.py``; registering them costs nothing when the ops are not used.

    torch.ops.modfx.flanger_chorus(x, mod_lo, feedback, min_delay_width, width, depth, mix, m_min, m_lfo)
    torch.ops.modfx.interp_linear(x, n, align_corners)
    torch.ops.modfx.lfo(freq, phase, shape, exp, n, sr)
    torch.ops.modfx.phaser(x, rate_hz, depth, centre_hz, feedback, mix, sr, block)
    torch.ops.modfx.cnn_layernorm(x, x_is_nchw, eps, round_tf32)               # -> channels-last (B, H, W, C)
    torch.ops.modfx.cnn_conv_pool_prelu(x, weight, bias, slope, dil_w, tf32)    # tcgen05 when tf32
    torch.ops.modfx.cnn_head(x, weight, bias)                                   # -> (output, latent)
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _ops
from ._ops import ModSource

_lib = torch.library.Library("modfx", "DEF")
_lib.define("flanger_chorus(Tensor x, Tensor mod, Tensor feedback, Tensor min_delay_width, Tensor width, "
            "Tensor depth, Tensor mix, int m_min, int m_lfo) -> Tensor")
_lib.define("interp_linear(Tensor x, int n, bool align_corners) -> Tensor")
_lib.define("lfo(Tensor freq, Tensor phase, Tensor shape, Tensor exp, int n, float sr) -> Tensor")
_lib.define("phaser(Tensor x, Tensor rate_hz, Tensor depth, Tensor centre_hz, Tensor feedback, Tensor mix, "
            "float sr, int block) -> Tensor")
_lib.define("cnn_layernorm(Tensor x, bool x_is_nchw, float eps, bool round_tf32) -> Tensor")
_lib.define("cnn_conv_pool_prelu(Tensor x, Tensor weight, Tensor bias, Tensor slope, int dil_w, bool tf32) -> Tensor")
_lib.define("cnn_head(Tensor x, Tensor weight, Tensor bias) -> (Tensor, Tensor)")


def _flanger_chorus(x: Tensor, mod: Tensor, feedback: Tensor, min_delay_width: Tensor, width: Tensor, depth: Tensor,
                    mix: Tensor, m_min: int, m_lfo: int) -> Tensor:
    src = ModSource.audio_rate(mod) if mod.size(-1) == x.size(-1) else ModSource.control_rate(mod)
    return _ops.flanger_chorus(x, src, m_min, m_lfo, feedback, min_delay_width, width, depth, mix)


def _interp_linear(x: Tensor, n: int, align_corners: bool) -> Tensor:
    return _ops.interp_linear(x, n, align_corners)


def _lfo(freq: Tensor, phase: Tensor, shape: Tensor, exp: Tensor, n: int, sr: float) -> Tensor:
    return _ops.lfo(n, sr, freq, phase, shape, exp)


def _phaser(x: Tensor, rate_hz: Tensor, depth: Tensor, centre_hz: Tensor, feedback: Tensor, mix: Tensor, sr: float,
            block: int) -> Tensor:
    return _ops.phaser(x, sr, rate_hz, depth, centre_hz, feedback, mix, block)


_impl = torch.library.Library("modfx", "IMPL", "CUDA")
_impl.impl("flanger_chorus", _flanger_chorus)
_impl.impl("interp_linear", _interp_linear)
_impl.impl("lfo", _lfo)
_impl.impl("phaser", _phaser)
_impl.impl("cnn_layernorm", lambda x, x_is_nchw, eps, round_tf32: _ops.cnn_layernorm(x, x_is_nchw, eps, round_tf32))
_impl.impl("cnn_conv_pool_prelu", lambda x, weight, bias, slope, dil_w, tf32: _ops.cnn_conv_pool_prelu(x, weight, bias, slope,
                                                                                                      dil_w, tf32))
_impl.impl("cnn_head", lambda x, weight, bias: _ops.cnn_head(x, weight, bias))
