"""GPU phaser: drop-in for the DSP of ``PedalboardPhaserDataset.apply_pedalboard_phaser``
(reference mod_extraction/datasets.py:455-482), which calls ``pedalboard.Phaser`` (JUCE dsp::Phaser).

PARITY UNPINNED: ``pedalboard==0.7.3`` is a third-party C++ wheel that is not available offline
(SURVEY F1).  The kernels follow this repo's own restatement of the JUCE algorithm and are
checked against its CPU version only; results must not be reported as "matches the reference".
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _ops, util

__all__ = ["Phaser", "apply_pedalboard_phaser", "PhaserRenderStep"]


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



class PhaserRenderStep:
    """GPU form of one BATCH of ``PedalboardPhaserDataset.__getitem__`` (datasets.py:428-453), the online phaser of
    ``train_lfo_phaser`` / ``interwoven_idmt_all``:

    * per example, in the reference's order: ``rate_hz`` (scipy log-uniform on numpy's global RNG), [the audio chunk of
      ``proc_n = n_samples + round(sr / rate_hz)`` samples -- supplied by the caller], ``depth`` (torch uniform),
      ``centre_frequency_hz`` (numpy log-uniform), ``feedback``, ``mix`` (torch uniform), then the crop position
      ``start_idx = randint(0, proc_n - n_samples + 1)`` (torch).  The four torch draws of an example are one generator
      word each, so the whole batch is drawn in two bulk calls with the same streams (``_rng.TorchMT``);
    * the phaser runs over each row from its first sample, and only the window ``[start_idx, start_idx + n_samples)``
      of wet and dry is produced (the effect is causal: nothing after the window is computed);
    * the ground-truth LFO ``make_mod_signal(proc_n, sr, rate_hz, pi / 2, "cos")[start_idx:...]`` resampled to
      ``n_samples // 100`` points comes from the closed form of the LFO (datasets.py:442-450).

    ``sample_params(B)`` -> dict; ``__call__(audio (B, 1, L), params)`` -> ``(dry, wet, mod_sig, fx_params)`` with
    ``L >= max_proc_n_samples`` (rows are read up to their own ``proc_n``).  PARITY UNPINNED like everything P1.
    """

    def __init__(self, fx_config: Dict[str, Dict[str, float]], n_samples: int, sr: float, buffer_size: int = 8192) -> None:
        self.cfg = fx_config["pedalboard_phaser"] if "pedalboard_phaser" in fx_config else fx_config
        self.n_samples, self.sr, self.buffer_size = n_samples, sr, buffer_size
        # datasets.py:420-421: the slowest LFO decides how long a chunk must be
        self.max_proc_n_samples = n_samples + int((sr / self.cfg["rate_hz"]["min"]) + 0.5)

    def sample_params(self, batch_size: int) -> Dict[str, object]:
        import numpy as np
        from scipy.stats import loguniform
        from ._rng import TorchMT, words_to_uniform
        c = self.cfg
        u = np.random.uniform(size=2 * batch_size)                      # numpy stream: rate, centre, rate, centre, ...
        lu = lambda q, r: (np.full_like(q, r["min"]) if r["min"] == r["max"] else loguniform.ppf(q, r["min"], r["max"]))
        rate = lu(u[0::2], c["rate_hz"])
        centre = lu(u[1::2], c["centre_frequency_hz"])
        mt = TorchMT()
        w = mt.words(4 * batch_size).reshape(batch_size, 4)             # torch stream: depth, feedback, mix, start_idx
        mt.consume(4 * batch_size)
        rate_n = (self.sr / rate + 0.5).astype(np.int64)                # datasets.py:433
        return {"rate_hz": rate, "centre_frequency_hz": centre,
                "depth": words_to_uniform(w[:, 0], c["depth"]["min"], c["depth"]["max"]),
                "feedback": words_to_uniform(w[:, 1], c["feedback"]["min"], c["feedback"]["max"]),
                "mix": words_to_uniform(w[:, 2], c["mix"]["min"], c["mix"]["max"]),
                "proc_n_samples": self.n_samples + rate_n,
                "start_idx": (w[:, 3] % (rate_n + 1).astype(np.uint32)).astype(np.int64)}     # datasets.py:445

    @torch.no_grad()
    def __call__(self, audio: Tensor, params: Optional[Dict[str, object]] = None):
        assert audio.ndim == 3 and audio.size(1) == 1, "mono (B, 1, L)"
        B, _, L = audio.shape
        p = self.sample_params(B) if params is None else params
        assert int(p["proc_n_samples"].max()) <= L, "rows shorter than n_samples + one LFO period (datasets.py:434-436)"
        on_cpu = not audio.is_cuda
        if on_cpu and not torch.cuda.is_available():
            raise RuntimeError("mod_extraction_b200 needs a CUDA device (no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device()) if on_cpu else audio.device
        x = (audio.pin_memory() if on_cpu and not audio.is_pinned() else audio).to(dev, non_blocking=True).float()
        t32 = lambda a: torch.from_numpy(a.astype("float32")).to(dev, non_blocking=True)
        start = torch.from_numpy(p["start_idx"].astype("int32")).to(dev, non_blocking=True)
        rate = t32(p["rate_hz"])
        wet, dry = _ops.phaser_crop(x.reshape(B, L), self.n_samples, start, self.sr, rate, t32(p["depth"]),
                                    t32(p["centre_frequency_hz"]), t32(p["feedback"]), t32(p["mix"]), self.buffer_size)
        from ._lib import SHAPE_ID
        mod_sig = _ops.lfo_window(self.n_samples // 100, self.n_samples, self.sr, rate,
                                  torch.full((B,), math.pi / 2, device=dev),
                                  torch.full((B,), SHAPE_ID["cos"], dtype=torch.int32, device=dev), start)
        fx_params = {"depth": torch.from_numpy(p["depth"]), "feedback": torch.from_numpy(p["feedback"]),
                     "mix": torch.from_numpy(p["mix"]), "rate_hz": torch.from_numpy(p["rate_hz"]), "shape": ["cos"] * B}
        dry, wet = dry.unsqueeze(1), wet.unsqueeze(1)
        if on_cpu:
            dry, wet, mod_sig = dry.cpu(), wet.cpu(), mod_sig.cpu()
        return dry, wet, mod_sig, fx_params
