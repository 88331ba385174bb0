"""CPU oracle for the effect-rendering hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package ``mod_extraction_b200`` never does.

It restates, on the CPU, the arithmetic of christhetree/mod_extraction's

* ``fx.py``            flanger/chorus delay line and tremolo       (C, modfx_oracle.c)
* ``modulations.py``   LFO shapes, corner finding, quasi-periodic
                       and combined LFOs                           (C + numpy)
* ``util.py``          linear_interpolate_last_dim                  (C)
* ``models.py``        log-mel front end of Spectral2DCNN           (numpy)
* ``datasets.py:455``  pedalboard/JUCE phaser                       (C, PARITY UNPINNED)

The float32 routines are pinned against golden vectors produced by the
reference itself (tests/golden/make_golden.py, run where /root/reference is
importable).  The phaser has no pin: pedalboard==0.7.3 is not available offline.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmodfx_oracle.so")

SHAPES = ["cos", "rect_cos", "inv_rect_cos", "tri", "saw", "rsaw", "sqr"]
SHAPE_ID = {s: i for i, s in enumerate(SHAPES)}

Param = Union[float, np.ndarray]


def build(force: bool = False) -> str:
    """Compile modfx_oracle.c with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "modfx_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        i64 = ctypes.c_int64
        ci = ctypes.c_int
        L.modfx_oracle_flanger_chorus.argtypes = [fp, fp, ci, fp, ci, ci, i64, ci, ci] + [fp] * 6
        L.modfx_oracle_flanger_chorus.restype = None
        L.modfx_oracle_flanger_chorus_allpass.argtypes = L.modfx_oracle_flanger_chorus.argtypes
        L.modfx_oracle_flanger_chorus_allpass.restype = None
        L.modfx_oracle_tremolo.argtypes = [fp, fp, ci, fp, ci, ci, i64, fp, fp]
        L.modfx_oracle_tremolo.restype = None
        L.modfx_oracle_lfo.argtypes = [fp, i64, ctypes.c_float, ctypes.c_float, ctypes.c_float, ci,
                                       ctypes.c_double]
        L.modfx_oracle_lfo.restype = None
        L.modfx_oracle_interp_linear.argtypes = [fp, fp, i64, i64, i64, ci]
        L.modfx_oracle_interp_linear.restype = None
        L.modfx_oracle_phaser.argtypes = [fp, fp, ci, i64, ctypes.c_float] + [fp] * 5 + [ci]
        L.modfx_oracle_phaser.restype = None
        L.modfx_oracle_num_threads.restype = ci
        L.modfx_oracle_set_threads.argtypes = [ci]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().modfx_oracle_num_threads())


def set_threads(n: int) -> None:
    lib().modfx_oracle_set_threads(int(n))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------------- fx.py

def delay_samples(sr: float, max_min_delay_ms: float, max_lfo_delay_ms: float) -> Tuple[int, int]:
    """fx.py:40-41."""
    m_min = int(((max_min_delay_ms / 1000.0) * sr) + 0.5)
    m_lfo = int(((max_lfo_delay_ms / 1000.0) * sr) + 0.5)
    return m_min, m_lfo


def _is_arr(p) -> bool:
    return isinstance(p, np.ndarray)


def derived_fc_params(B: int, m_min: int, m_lfo: int, feedback: Param, min_delay_width: Param,
                      width: Param, depth: Param, mix: Param):
    """Per-example float32 coefficients with torch's promotion rules (fx.py:97-98,114-117).

    A (B,) float32 array behaves like a torch tensor (float32 arithmetic with the
    python int converted to float32); a python float behaves like a python scalar
    (double arithmetic between python numbers, rounded to float32 when it meets a tensor).
    """
    one = np.float32(1.0)

    def bc(v):
        return np.full((B,), v, dtype=np.float32)

    if _is_arr(width):
        lfo_delay = np.float32(m_lfo) * _f32(width)
    else:
        lfo_delay = bc(np.float32(float(m_lfo) * float(width)))
    if _is_arr(min_delay_width):
        min_delay = _f32(min_delay_width) * np.float32(m_min)
    else:
        min_delay = bc(np.float32(float(min_delay_width) * float(m_min)))
    fb = _f32(feedback) if _is_arr(feedback) else bc(np.float32(feedback))
    dp = _f32(depth) if _is_arr(depth) else bc(np.float32(depth))
    if _is_arr(mix):
        mx = _f32(mix)
        omm = one - mx
    else:
        mx = bc(np.float32(mix))
        omm = bc(np.float32(1.0 - float(mix)))
    return [_f32(a) for a in (lfo_delay, min_delay, fb, dp, mx, omm)]


def flanger_chorus(x: np.ndarray, mod_sig: np.ndarray, feedback: Param = 0.0,
                   min_delay_width: Param = 1.0, width: Param = 1.0, depth: Param = 1.0,
                   mix: Param = 1.0, *, sr: float = 44100.0, max_min_delay_ms: float,
                   max_lfo_delay_ms: float, interpolation: str = "linear") -> np.ndarray:
    """MonoFlangerChorusModule.forward, fx.py:121-130 (ctor fx.py:26-44).
    interpolation="allpass": the same delay line with a first-order all-pass fractional-delay interpolator -- OWN
    definition (the reference only has the linear one), see modfx_oracle.c."""
    assert interpolation in ("linear", "allpass")
    x = _f32(x)
    assert x.ndim == 3
    B, C, N = x.shape
    mod_sig = _f32(mod_sig)
    assert mod_sig.shape[0] == B and mod_sig.shape[-1] == N
    has_ch = 1 if mod_sig.ndim == 3 else 0
    m_min, m_lfo = delay_samples(sr, max_min_delay_ms, max_lfo_delay_ms)
    coefs = derived_fc_params(B, m_min, m_lfo, feedback, min_delay_width, width, depth, mix)
    y = np.empty_like(x)
    fn = lib().modfx_oracle_flanger_chorus if interpolation == "linear" else lib().modfx_oracle_flanger_chorus_allpass
    fn(_ptr(x), _ptr(mod_sig), has_ch, _ptr(y), B, C, N, m_min, m_lfo, *[_ptr(c) for c in coefs])
    return y


def flanger_chorus_allpass_f64(x: np.ndarray, mod_sig: np.ndarray, feedback: float, min_delay_width: float, width: float,
                               depth: float, mix: float, *, sr: float = 44100.0, max_min_delay_ms: float,
                               max_lfo_delay_ms: float) -> np.ndarray:
    """The all-pass variant in float64, one (N,) clip, scalar parameters, python loop (small cases only): the
    "what it should compute" restatement of this repository's OWN definition."""
    x = np.asarray(x, dtype=np.float64)
    mod = np.asarray(mod_sig, dtype=np.float64)
    m_min, m_lfo = delay_samples(sr, max_min_delay_ms, max_lfo_delay_ms)
    M = m_min + m_lfo
    buf = np.zeros(M)
    y = np.empty_like(x)
    it_prev = 0.0
    for n in range(x.shape[0]):
        w = n % M
        d = (m_lfo * width) * mod[n] + min_delay_width * m_min
        r = ((w - d) + M) % M
        p = int(math.floor(r))
        fr = r - p
        p = min(max(p, 0), M - 1)
        q = (p + 1) % M
        it = fr / (2.0 - fr) * (buf[q] - it_prev) + buf[p]
        it_prev = it
        buf[w] = x[n] + feedback * it
        y[n] = min(max((1.0 - mix) * x[n] + mix * (x[n] + depth * it), -1.0), 1.0)
    return y


def tremolo(x: np.ndarray, mod_sig: np.ndarray, mix: Param = 1.0) -> np.ndarray:
    """apply_tremolo, fx.py:13-22."""
    x = _f32(x)
    B, C, N = x.shape
    mod_sig = _f32(mod_sig)
    has_ch = 1 if mod_sig.ndim == 3 else 0
    if _is_arr(mix):
        mx = _f32(mix)
        omm = _f32(np.float32(1.0) - mx)
    else:
        mx = np.full((B,), np.float32(mix), dtype=np.float32)
        omm = np.full((B,), np.float32(1.0 - float(mix)), dtype=np.float32)
    y = np.empty_like(x)
    lib().modfx_oracle_tremolo(_ptr(x), _ptr(mod_sig), has_ch, _ptr(y), B, C, N, _ptr(mx), _ptr(omm))
    return y


# --------------------------------------------------------------------------- modulations.py

def make_mod_signal(n_samples: int, sr: float, freq: float, phase: float = 0.0, shape: str = "cos",
                    exp: float = 1.0) -> np.ndarray:
    """make_mod_signal, modulations.py:16-57 (asserts :22-30 included)."""
    assert n_samples > 0
    assert 0.0 < freq < sr / 2.0
    assert -2 * math.pi <= phase <= 2 * math.pi
    assert shape in SHAPE_ID
    freq = float(freq)
    phase = float(phase)
    if shape in ("rect_cos", "inv_rect_cos"):
        freq /= 2.0
        phase /= 2.0
    assert exp > 0
    out = np.empty((n_samples,), dtype=np.float32)
    lib().modfx_oracle_lfo(_ptr(out), n_samples, np.float32(sr), np.float32(freq), np.float32(phase),
                           SHAPE_ID[shape], float(exp))
    return out


def linear_interpolate_last_dim(x: np.ndarray, n: int, align_corners: bool = True) -> np.ndarray:
    """util.linear_interpolate_last_dim, util.py:15-29."""
    x = _f32(x)
    assert 1 <= x.ndim <= 3
    if x.shape[-1] == n:
        return x
    rows = int(np.prod(x.shape[:-1])) if x.ndim > 1 else 1
    out = np.empty(x.shape[:-1] + (n,), dtype=np.float32)
    lib().modfx_oracle_interp_linear(_ptr(x), _ptr(out), rows, x.shape[-1], n, 1 if align_corners else 0)
    return out


def find_corners(mod_sig: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """find_corners, modulations.py:219-238 (float32 arithmetic kept)."""
    m = _f32(mod_sig)
    assert m.ndim == 2
    diff = m[:, 1:] - m[:, :-1]
    diff_r = diff[:, 1:]
    diff_l = diff[:, :-1]
    zero = np.float32(0.0)
    diff_pos_l = np.where(diff_l > 0, diff_l, zero).astype(np.float32)
    diff_neg_l = np.where(diff_l < 0, diff_l, zero).astype(np.float32)
    dr = (diff_r + np.float32(1e-16)).astype(np.float32)
    top = -np.floor((diff_pos_l * dr).astype(np.float32)).astype(np.int64)
    bot = -np.floor((diff_neg_l * dr).astype(np.float32)).astype(np.int64)
    top_c = np.zeros_like(m)
    bot_c = np.zeros_like(m)
    top_c[:, 1:-1] = top
    bot_c[:, 1:-1] = bot
    return top_c, bot_c


class ReplayDraws:
    """Replays the host RNG draws the reference made (util.sample_uniform / util.choice)."""

    def __init__(self, uniforms: Sequence[float] = (), choices: Sequence[int] = ()):
        self.u = list(uniforms)
        self.c = list(choices)
        self.ui = 0
        self.ci = 0

    def uniform(self, low: float, high: float) -> float:
        """util.sample_uniform (util.py:45-49): float32 tensor math, returned as python float.
        The replayed value is the raw torch.rand() draw."""
        r = np.float32(self.u[self.ui])
        self.ui += 1
        return float(np.float32(np.float32(r * np.float32(high - low)) + np.float32(low)))

    def choice(self, n: int) -> int:
        """util.choice (util.py:32-35): index drawn with torch.randint."""
        v = int(self.c[self.ci])
        self.ci += 1
        assert 0 <= v < n
        return v


def _time_stretch_len(size: int, rng, l_min, l_max, r_min, r_max, lr_split) -> int:
    """_time_stretch_section, modulations.py:104-118 (length decision only)."""
    if rng.uniform(0.0, 1.0) < lr_split:
        xx = int((rng.uniform(l_min, l_max) * size) + 0.5)
        return max(2, size - xx)
    xx = int((rng.uniform(r_min, r_max) * size) + 0.5)
    return size + xx


def make_quasi_periodic(mod_sig: np.ndarray, l_min: float = 0.2, l_max: float = 0.2, r_min: float = 0.2,
                        r_max: float = 0.2, lr_split: float = 0.5, *, rng) -> np.ndarray:
    """make_quasi_periodic, modulations.py:121-160."""
    m = _f32(mod_sig)
    assert m.ndim == 1
    top, bot = find_corners(m[None, :])
    corners = top if top.sum() > bot.sum() else bot
    idxs = [int(i) for i in np.nonzero(corners[0] == 1)[0]]
    if len(idxs) < 2:
        return m
    prev = 0
    sections: List[np.ndarray] = []
    total = 0
    for idx in idxs:
        sec = m[prev:idx + 1]
        new_len = _time_stretch_len(sec.shape[0], rng, l_min, l_max, r_min, r_max, lr_split)
        new_sec = linear_interpolate_last_dim(sec, new_len, True)[:-1]
        total += new_sec.shape[0]
        sections.append(new_sec)
        prev = idx
    orig = m.shape[0]
    tail = m[prev:orig]
    total += tail.shape[0]
    if total < orig:
        tail = linear_interpolate_last_dim(tail, tail.shape[0] + (orig - total), True)
    sections.append(tail)
    return np.concatenate(sections)[:orig].astype(np.float32)


def make_combined_mod_sig(n_samples: int, sr: float, freq: float, phase: float, shapes: List[str], *,
                          rng) -> np.ndarray:
    """make_combined_mod_sig, modulations.py:191-210."""
    cur = shapes[rng.choice(len(shapes))]
    m = make_mod_signal(n_samples, sr, freq, phase, cur).copy()
    _, bot = find_corners(m[None, :])
    idxs = [int(i) for i in np.nonzero(bot[0] == 1)[0]]
    if len(idxs) > 1:
        for i, idx in enumerate(idxs[1:]):
            prev = idxs[i]
            sec_len = idx - prev + 1
            cur = shapes[rng.choice(len(shapes))]
            m[prev:idx + 1] = make_mod_signal(sec_len, sec_len, 1.0, 0.0, cur)
    return m


# --------------------------------------------------------------------------- models.py (log-mel)

def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace float32 CPU: symmetric halves (RangeFactories.cpp)."""
    start32, end32 = np.float32(start), np.float32(end)
    step = np.float32((end32 - start32) / np.float32(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    lo = (start32 + step * i.astype(np.float32)).astype(np.float32)
    hi = (end32 - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, lo, hi).astype(np.float32)


def mel_filterbank(sr: int = 44100, n_fft: int = 1024, n_mels: int = 256, f_min: float = 0.0,
                   f_max: Optional[float] = None) -> np.ndarray:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') as used by
    MelSpectrogram in models.py:170-175.  Returns (n_fft//2+1, n_mels) float32."""
    n_freqs = n_fft // 2 + 1
    f_max = float(sr // 2) if f_max is None else f_max
    all_freqs = _linspace_f32(0.0, float(sr // 2), n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = _linspace_f32(m_min, m_max, n_mels + 2)
    f_pts = (np.float32(700.0) * (np.power(np.float32(10.0), m_pts / np.float32(2595.0), dtype=np.float32)
                                  - np.float32(1.0))).astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(np.float32(0.0), np.minimum(down, up)).astype(np.float32)
    return fb


def hann_periodic(n: int) -> np.ndarray:
    """torch.hann_window(n, periodic=True) in float32."""
    k = np.arange(n, dtype=np.float32)
    ang = (k * np.float32(2.0 * math.pi / n)).astype(np.float32)
    return (np.float32(0.5) - np.float32(0.5) * np.cos(ang, dtype=np.float32)).astype(np.float32)


def log_mel(x: np.ndarray, sr: int = 44100, n_fft: int = 1024, hop_len: int = 256, n_mels: int = 256,
            eps: float = 1e-7, fft_dtype=np.float32, fb: Optional[np.ndarray] = None) -> np.ndarray:
    """Spectral2DCNN front end: MelSpectrogram -> clip(min=eps) -> log (models.py:170-175,199,207-208).

    (..., T) -> (..., n_mels, T // hop_len + 1).  center=True reflect padding, periodic Hann,
    power 2.0, HTK mel, no normalisation.

    ``fb``: optional (n_fft//2+1, n_mels) filterbank.  torchaudio builds its table in float32
    and the triangle weights are ill-conditioned there (a 1-ulp change of a band edge near
    20 kHz moves a weight by ~3e-5, i.e. ~3e-4 in log-mel), so bit-level agreement with the
    reference needs the reference's own table; tests pass the golden copy of it.
    """
    x = _f32(x)
    lead = x.shape[:-1]
    T = x.shape[-1]
    xs = x.reshape(-1, T)
    pad = n_fft // 2
    assert T > pad
    xp = np.pad(xs, ((0, 0), (pad, pad)), mode="reflect")
    n_frames = T // hop_len + 1
    win = hann_periodic(n_fft)
    fb = mel_filterbank(sr, n_fft, n_mels) if fb is None else _f32(fb)
    idx = np.arange(n_frames)[:, None] * hop_len + np.arange(n_fft)[None, :]
    out = np.empty((xs.shape[0], n_mels, n_frames), dtype=np.float32)
    for r in range(xs.shape[0]):
        frames = (xp[r][idx] * win[None, :]).astype(fft_dtype)
        spec = np.fft.rfft(frames, axis=-1)
        power = (spec.real.astype(fft_dtype) ** 2 + spec.imag.astype(fft_dtype) ** 2).astype(np.float32)
        mel = power @ fb                                   # (frames, n_mels)
        mel = np.maximum(mel, np.float32(eps))
        out[r] = np.log(mel).T.astype(np.float32)
    return out.reshape(lead + (n_mels, n_frames))


# --------------------------------------------------------------------------- phaser (unpinned)

def phaser(x: np.ndarray, sr: float, rate_hz, depth, centre_frequency_hz, feedback, mix,
           block: int = 8192) -> np.ndarray:
    """Own restatement of pedalboard.Phaser -> juce::dsp::Phaser (datasets.py:455-482).
    PARITY UNPINNED: the dependency is not available offline.  x: (B, N) float32."""
    x = _f32(x)
    assert x.ndim == 2
    B, N = x.shape

    def per_ex(v):
        return _f32(np.broadcast_to(np.asarray(v, dtype=np.float32), (B,)))

    args = [per_ex(v) for v in (rate_hz, depth, centre_frequency_hz, feedback, mix)]
    y = np.empty_like(x)
    lib().modfx_oracle_phaser(_ptr(x), _ptr(y), B, N, np.float32(sr), *[_ptr(a) for a in args], int(block))
    return y


# --------------------------------------------------------------------------- extracted-LFO post-processing (N4)

def smoothen(x: np.ndarray, smooth_n_frames: int) -> np.ndarray:
    """smoothen, modulations.py:358-362: moving average over `smooth_n_frames` frames, no padding.

    torch's CPU mean over the last dim of the unfold view adds float32 in a fixed order, restated here: four
    8-lane vector accumulators over blocks of 32 elements, remaining whole 8-vectors into accumulators 0, 1, 2,
    the four accumulators added left to right, the 8 lanes added left to right, then the scalar tail; finally
    a true division by the window length.  Bitwise for the windows the reference ships (4, 8, and the default
    32) and every window that is < 5 or a multiple of 8; within 2e-7 otherwise."""
    x = _f32(x)
    w = int(smooth_n_frames)
    if w <= 1:
        return x
    win = np.lib.stride_tricks.sliding_window_view(x, w, axis=-1)          # (..., n_out, w)
    acc = np.zeros(win.shape[:-1] + (4, 8), dtype=np.float32)
    i = 0
    while i + 32 <= w:
        acc = (acc + win[..., i:i + 32].reshape(win.shape[:-1] + (4, 8))).astype(np.float32)
        i += 32
    j = 0
    while i + 8 <= w:
        acc[..., j, :] = (acc[..., j, :] + win[..., i:i + 8]).astype(np.float32)
        i += 8
        j += 1
    v = acc[..., 0, :]
    for j in range(1, 4):
        v = (v + acc[..., j, :]).astype(np.float32)
    s = v[..., 0]
    for k in range(1, 8):
        s = (s + v[..., k]).astype(np.float32)
    for k in range(i, w):
        s = (s + win[..., k]).astype(np.float32)
    return (s / np.float32(w)).astype(np.float32)


def _stretch_corners_1d(m: np.ndarray, top: np.ndarray, bottom: np.ndarray, top_val: float = 1.0,
                        bot_val: float = 0.0) -> np.ndarray:
    """_stretch_corners, modulations.py:259-291, float32 operation by operation.  Anchors that came from a
    tensor element (first / last sample) are float32, the corner targets are python floats; mixed arithmetic
    rounds exactly like torch's scalar promotion (float / tensor is reciprocal(tensor) * float)."""
    f32 = np.float32
    n = m.shape[0]
    anchors = [(int(i), f32(top_val)) for i in np.nonzero(top == 1)[0]] + \
              [(int(i), f32(bot_val)) for i in np.nonzero(bottom == 1)[0]] + [(n - 1, m[-1])]
    anchors.sort(key=lambda a: a[0])
    out = m.copy()
    prev_idx, prev_anchor = 0, m[0]
    for idx, target in anchors:
        seg = out[prev_idx + 1:idx + 1]
        if prev_anchor != target:
            curr_range = f32(abs(f32(m[prev_idx] - m[idx])))
            target_range = f32(abs(f32(prev_anchor - target)))
            with np.errstate(divide="ignore", invalid="ignore"):
                scale = f32(target_range / curr_range)
                seg -= seg.min() if seg.size else f32(0)
                seg *= scale
                if seg.size:
                    seg += f32(target - seg[-1])
        prev_idx, prev_anchor = idx, target
    return out


def stretch_corners(mod_sig: np.ndarray, max_n_corners: int = 10, smooth_n_frames: int = 32) -> np.ndarray:
    """stretch_corners, modulations.py:294-307."""
    m = smoothen(_f32(mod_sig), smooth_n_frames)
    assert m.ndim == 2
    top, bottom = find_corners(m)
    rows = []
    for r in range(m.shape[0]):
        if top[r].sum() + bottom[r].sum() > max_n_corners:
            rows.append(m[r])
        else:
            rows.append(_stretch_corners_1d(m[r], top[r], bottom[r]))
    return np.stack(rows, axis=0)


def check_mod_sig(top: np.ndarray, bottom: np.ndarray, min_top_corners: int = 1, max_top_corners: int = 6,
                  min_bottom_corners: int = 1, max_bottom_corners: int = 6,
                  min_fraction_between_corners: float = 0.10) -> bool:
    """check_mod_sig, modulations.py:311-345 (the signal itself only contributes its length)."""
    n_top, n_bot = int(top.sum()), int(bottom.sum())
    if n_top < min_top_corners or n_bot < min_bottom_corners:
        return False
    if n_top > max_top_corners or n_bot > max_bottom_corners:
        return False
    min_n_frames = int(min_fraction_between_corners * top.shape[0])
    for c in (top, bottom):
        idx = np.nonzero(c == 1)[0]
        if idx.size > 1 and int(np.diff(idx).min()) < min_n_frames:
            return False
    return True


def find_valid_mod_sig_indices(mod_sig: np.ndarray) -> List[int]:
    """find_valid_mod_sig_indices, modulations.py:348-355."""
    m = _f32(mod_sig)
    top, bottom = find_corners(m)
    return [r for r in range(m.shape[0]) if check_mod_sig(top[r], bottom[r])]


# ---- LFO-net body, models.py:183-195,209-214 (numpy float32) ----------------------------
def layer_norm_2d(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.LayerNorm([H, W], elementwise_affine=False), models.py:186: per (b, c) over the last two dims,
    biased variance.  x (B, C, H, W)."""
    x = _f32(x)
    mean = x.mean(axis=(-2, -1), keepdims=True, dtype=np.float64)
    var = ((x.astype(np.float64) - mean) ** 2).mean(axis=(-2, -1), keepdims=True)
    return ((x - mean) / np.sqrt(var + eps)).astype(np.float32)


def conv2d_same(x: np.ndarray, weight: np.ndarray, bias: np.ndarray, dil_w: int = 1, tf32: bool = False) -> np.ndarray:
    """nn.Conv2d(Cin, Cout, (KH, KW), dilation=(1, dil_w), padding="same"), models.py:187.
    x (B, Cin, H, W), weight (Cout, Cin, KH, KW).  One float32 GEMM per kernel tap; ``tf32`` rounds both
    operands to 10 mantissa bits first (what the tensor-core path feeds its MMAs)."""
    x, weight = _f32(x), _f32(weight)
    if tf32:
        x, weight = round_tf32(x), round_tf32(weight)
    B, Cin, H, W = x.shape
    Cout, _, KH, KW = weight.shape
    ph, pw = KH // 2, (KW // 2) * dil_w
    xp = np.zeros((B, Cin, H + 2 * ph, W + 2 * pw), dtype=np.float32)
    xp[:, :, ph:ph + H, pw:pw + W] = x
    out = np.zeros((B, Cout, H, W), dtype=np.float32)
    for kh in range(KH):
        for kw in range(KW):
            win = xp[:, :, kh:kh + H, kw * dil_w:kw * dil_w + W]              # (B, Cin, H, W)
            out += np.einsum("oc,bchw->bohw", weight[:, :, kh, kw], win, optimize=True).astype(np.float32)
    return out + _f32(bias)[None, :, None, None]


def round_tf32(a: np.ndarray) -> np.ndarray:
    """float32 -> nearest TF32 (ties away from zero, cvt.rna.tf32.f32)."""
    bits = _f32(a).view(np.int32)
    return ((bits + 0x1000) & ~0x1FFF).astype(np.int32).view(np.float32)


def max_pool_h2(x: np.ndarray) -> np.ndarray:
    """nn.MaxPool2d((2, 1)), models.py:188."""
    B, C, H, W = x.shape
    return x[:, :, :H - (H % 2)].reshape(B, C, H // 2, 2, W).max(axis=3)


def prelu(x: np.ndarray, slope: np.ndarray) -> np.ndarray:
    """nn.PReLU(num_parameters=C), models.py:189."""
    return np.where(x > 0, x, _f32(slope)[None, :, None, None] * x).astype(np.float32)


def spectral_2dcnn_body(logmel: np.ndarray, convs, out_w: np.ndarray, out_b: np.ndarray, temp_dilations,
                        ln_eps: float = 1e-5, tf32_from_layer: Optional[int] = None):
    """Spectral2DCNN.forward after clip + log, models.py:209-214.
    logmel (B, C, n_mels, n_frames); convs = [(weight (Cout, Cin, KH, KW), bias, prelu slope), ...];
    out_w (L, C), out_b (L,).  Returns (sigmoid output (B, L, W), latent (B, C, W)).
    ``tf32_from_layer`` = index of the first layer whose operands are TF32-rounded (None: float32 all the way)."""
    x = _f32(logmel)
    for i, ((w, b, a), d) in enumerate(zip(convs, temp_dilations)):
        x = layer_norm_2d(x, ln_eps)
        x = conv2d_same(x, w, b, d, tf32=tf32_from_layer is not None and i >= tf32_from_layer)
        x = prelu(max_pool_h2(x), a)
    latent = x.mean(axis=-2, dtype=np.float32)
    z = np.einsum("lc,bcw->blw", _f32(out_w), latent).astype(np.float32) + _f32(out_b)[None, :, None]
    return (1.0 / (1.0 + np.exp(-z))).astype(np.float32), latent
