"""Log-mel front end of the reference's LFO extractor, on the GPU.

Reference: ``Spectral2DCNN`` builds ``self.spectrogram = MelSpectrogram(sample_rate=int(sr), n_fft,
hop_length, normalized=False, n_mels, center=True)`` (models.py:170-175) and its ``forward`` does
``spectrogram(x)`` -> [SpecAugment when training] -> ``clip(min=eps)`` -> ``log`` (models.py:199-208).
``LogMelLoss`` (losses.py:105-130) uses the same front end.

``MelSpectrogram`` here is a drop-in for that ``spectrogram`` attribute (same constructor keywords,
returns mel *power*, (..., T) -> (..., n_mels, T // hop + 1)); ``LogMelSpectrogram`` fuses the clip
and log into the same kernel for inference, where no masking sits in between.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib

__all__ = ["MelSpectrogram", "LogMelSpectrogram", "mel_filterbank", "banded"]


def _hz_to_mel(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, f_min: float = 0.0, f_max: Optional[float] = None) -> Tensor:
    """HTK triangular filterbank, norm=None: (n_fft//2+1, n_mels) float32.

    torchaudio builds this table in float32, where the triangle edges are ill-conditioned (one ulp
    of a band edge near 20 kHz moves a weight by ~3e-5, i.e. ~3e-4 in log-mel).  To be a bit-level
    drop-in we use torchaudio's own generator when it is importable and fall back to a float64
    evaluation (closer to the mathematical filterbank, up to ~3e-5 away from torchaudio's)."""
    f_max = float(sr // 2) if f_max is None else f_max
    n_freqs = n_fft // 2 + 1
    try:
        from torchaudio.functional import melscale_fbanks
        return melscale_fbanks(n_freqs, f_min, f_max, n_mels, sr, norm=None, mel_scale="htk").float()
    except Exception:       # torchaudio missing: mathematically exact table
        all_freqs = np.linspace(0.0, sr // 2, n_freqs)
        m_pts = np.linspace(_hz_to_mel(f_min), _hz_to_mel(f_max), n_mels + 2)
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts[None, :] - all_freqs[:, None]
        down = -slopes[:, :-2] / f_diff[:-1]
        up = slopes[:, 2:] / f_diff[1:]
        return torch.from_numpy(np.maximum(0.0, np.minimum(down, up)).astype(np.float32))


def banded(fb: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """(n_freqs, n_mels) filterbank -> (start (n_mels,), count (n_mels,), weights (n_mels, stride)).
    The table is 0.77 % dense (<= 14 taps per band for the reference configuration)."""
    fbn = fb.detach().cpu().numpy()
    n_freqs, n_mels = fbn.shape
    start = np.zeros(n_mels, dtype=np.int32)
    count = np.zeros(n_mels, dtype=np.int32)
    for m in range(n_mels):
        nz = np.nonzero(fbn[:, m])[0]
        if nz.size:
            start[m] = nz[0]
            count[m] = nz[-1] - nz[0] + 1
    stride = max(4, int(-(-int(count.max()) // 4) * 4))
    w = np.zeros((n_mels, stride), dtype=np.float32)
    for m in range(n_mels):
        w[m, :count[m]] = fbn[start[m]:start[m] + count[m], m]
    return torch.from_numpy(start), torch.from_numpy(count), torch.from_numpy(w)


class MelSpectrogram(nn.Module):
    """GPU replacement for the ``spectrogram`` attribute of Spectral2DCNN / LogMelLoss."""

    apply_log = False

    def __init__(self, sample_rate: int = 44100, n_fft: int = 1024, hop_length: int = 256, n_mels: int = 256,
                 normalized: bool = False, center: bool = True, eps: float = 1e-7,
                 fb: Optional[Tensor] = None, window: Optional[Tensor] = None) -> None:
        super().__init__()
        assert not normalized and center, "only the reference configuration (normalized=False, center=True)"
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels, self.eps = sample_rate, n_fft, hop_length, n_mels, eps
        fb = mel_filterbank(sample_rate, n_fft, n_mels) if fb is None else fb.detach().float().cpu()
        assert fb.shape == (n_fft // 2 + 1, n_mels)
        window = torch.hann_window(n_fft, periodic=True) if window is None else window.detach().float().cpu()
        start, count, weight = banded(fb)
        self.fb_taps = max(1, int(count.sum()))
        self.register_buffer("fb", fb, persistent=False)
        self.register_buffer("window", window, persistent=False)
        self.register_buffer("fb_start", start, persistent=False)
        self.register_buffer("fb_count", count, persistent=False)
        self.register_buffer("fb_weight", weight, persistent=False)

    @classmethod
    def from_torchaudio(cls, mel, eps: float = 1e-7):
        """Build from an existing torchaudio.transforms.MelSpectrogram (e.g. ``net.spectrogram``):
        reuses its very filterbank and window buffers."""
        return cls(sample_rate=mel.sample_rate, n_fft=mel.n_fft, hop_length=mel.hop_length, n_mels=mel.n_mels,
                   eps=eps, fb=mel.mel_scale.fb, window=mel.spectrogram.window)

    def n_frames(self, n_samples: int) -> int:
        return n_samples // self.hop_length + 1

    @torch.no_grad()
    def forward_rows(self, x: Tensor, T: int, n_rows: int, out: Tensor, x_row_stride: int, out_row_stride: int,
                     row_index: Optional[Tensor] = None) -> None:
        """Strided / indexed variant used by the batch renderer: row r of the input starts at
        ``x.data_ptr() + 4*r*x_row_stride`` and its (n_mels, n_frames) result is written at
        ``out.data_ptr() + 4*r*out_row_stride``; ``row_index`` (int32, CUDA) selects rows."""
        if self.fb_weight.device != x.device:
            self.to(x.device)
        vp = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().modfx_logmel_f32(
                vp(x), vp(out), n_rows, T, self.n_fft, self.hop_length, self.n_mels, vp(self.window), vp(self.fb_start),
                vp(self.fb_count), vp(self.fb_weight), self.fb_weight.size(1), self.fb_taps, float(self.eps),
                1 if self.apply_log else 0, x_row_stride, out_row_stride, vp(row_index),
                0 if row_index is None else row_index.numel(),
                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    @torch.no_grad()
    def forward(self, x: Tensor, out: Optional[Tensor] = None) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("modfx: the log-mel front end needs a CUDA tensor (no CPU kernel)")
        if self.fb_weight.device != x.device:
            self.to(x.device)
        x = x.detach().float().contiguous()
        T = x.size(-1)
        lead = x.shape[:-1]
        R = x.numel() // T
        nf = self.n_frames(T)
        if out is None:
            out = torch.empty(lead + (self.n_mels, nf), device=x.device, dtype=torch.float32)
        assert out.is_contiguous() and out.numel() == R * self.n_mels * nf
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().modfx_logmel_f32(
                vp(x), vp(out), R, T, self.n_fft, self.hop_length, self.n_mels, vp(self.window), vp(self.fb_start),
                vp(self.fb_count), vp(self.fb_weight), self.fb_weight.size(1), self.fb_taps, float(self.eps),
                1 if self.apply_log else 0, 0, 0, None, 0,
                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out


class LogMelSpectrogram(MelSpectrogram):
    """``log(clip(MelSpectrogram(x), min=eps))`` in one kernel (models.py:199,207-208)."""

    apply_log = True
