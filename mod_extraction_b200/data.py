"""Callers and data formats on either side of the render path.

N1  the batched host step that feeds the renderer: per-example LFO parameter sampling of
    ``RandomAudioChunkAndModSigDataset.__getitem__`` (datasets.py:365-398) and the five effect-parameter draws
    of ``FlangerCPUDataModule.on_before_batch_transfer`` (data_modules.py:419-458), with the reference's RNG
    sources and draw order (torch global CPU generator; scipy/numpy global RNG for log-uniform rates), plus
    the ``fx_params`` dict layout the rest of the reference expects.
N2  the on-disk cache of pre-rendered examples that ``train_lfo_interwoven_all`` consumes
    (``<md5(fx_params)>.pt`` + ``<md5>_dry.wav`` + ``<md5>_wet.wav``): writer as in scripts/scratch.py:145-162,
    reader as ``PreprocessedDataset.__getitem__`` (datasets.py:504-535).  WAV files are 32-bit float like
    ``torchaudio.save`` writes for float tensors; they are written and read with scipy (torchaudio.save needs
    TorchCodec, which this image does not ship).

Everything here is host logic: the arithmetic stays in the kernels.
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch as tr
from torch import Tensor as T

from . import util
from .fx import MonoFlangerChorusModule
from .modulations import (SHAPE_ID, make_combined_mod_sig_batch, make_mod_signal_batch,
                          make_quasi_periodic_batch)

FLANGER_PARAM_ORDER = ("feedback", "min_delay_width", "width", "depth", "mix")     # data_modules.py:421-445


# --------------------------------------------------------------------------------------------- N1

def sample_mod_sig_batch(mod_cfg: Dict[str, Any], batch_size: int, n_samples: int, sr: float,
                         device=None, exact_stream: bool = True) -> Tuple[T, Dict[str, Any]]:
    """Batched LFO part of RandomAudioChunkAndModSigDataset.__getitem__ (datasets.py:367-397).

    Draw order per example, as in the reference: rate (scipy log-uniform on numpy's global RNG), phase (torch
    uniform), shape (torch randint), then -- BEFORE the next example's triple -- the draws of
    ``make_combined_mod_sig`` ("combined" configs: base shape + one shape per span) and of ``make_quasi_periodic``
    ("quasiperiodic" configs: two draws per section).  How many draws an example makes depends on the corners of
    its own signal, and its phase comes from the same stream, so for those configs the stream can only be followed
    example by example: ``exact_stream=True`` (default) does that (one small launch sequence per example, like the
    reference's per-item ``__getitem__``), and under a given seed reproduces the reference's LFOs and ``fx_params``.
    ``exact_stream=False`` draws all (rate, phase, shape) triples first and lets the batched generators make the
    section draws afterwards: one launch sequence per batch, same distribution, different stream for B > 1.
    Plain LFO configs need no such choice: the batched form IS the reference's stream.
    Returns (mod_sig (B, n_samples // 100) on the GPU, fx_params with rate_hz, phase, shape, exp).
    """
    n_lo, sr_lo = n_samples // 100, sr // 100                                   # datasets.py:377-382
    exp = mod_cfg["exp"]
    combined, quasi = bool(mod_cfg.get("combined")), bool(mod_cfg.get("quasiperiodic"))
    q_args = [mod_cfg.get(k, 0.2) for k in ("l_min", "l_max", "r_min", "r_max")] + [mod_cfg.get("lr_split", 0.5)]
    rates, phases, shapes, rows = [], [], [], []
    per_example = exact_stream and (combined or quasi) and batch_size > 1
    for _ in range(batch_size):
        rates.append(util.sample_log_uniform(mod_cfg["rate_hz"]["min"], mod_cfg["rate_hz"]["max"]))
        phases.append(util.sample_uniform(mod_cfg["phase"]["min"], mod_cfg["phase"]["max"]))
        shapes.append(util.choice(mod_cfg["shapes"]))
        if per_example:                                                         # datasets.py:375-390, one item
            if combined:
                row = make_combined_mod_sig_batch(n_lo, sr_lo, rates[-1:], phases[-1:], mod_cfg["shapes"], device)
            else:
                row = make_mod_signal_batch(n_lo, sr_lo, rates[-1:], phases[-1:], shapes[-1:],
                                            None if exp == 1.0 else [exp], device)
            if quasi:
                row = make_quasi_periodic_batch(row, *q_args)
            rows.append(row)
    if per_example:
        mod_sig = tr.cat(rows, 0)
    else:
        if combined:
            mod_sig = make_combined_mod_sig_batch(n_lo, sr_lo, rates, phases, mod_cfg["shapes"], device)
        else:
            mod_sig = make_mod_signal_batch(n_lo, sr_lo, rates, phases, shapes,
                                            None if exp == 1.0 else [exp] * batch_size, device)
        if quasi:
            mod_sig = make_quasi_periodic_batch(mod_sig, *q_args)
    fx_params = {"rate_hz": tr.tensor(rates, dtype=tr.float64), "phase": tr.tensor(phases, dtype=tr.float64),
                 "shape": shapes, "exp": tr.full((batch_size,), float(exp), dtype=tr.float64)}     # default_collate layout
    return mod_sig, fx_params


def sample_flanger_params(flanger_cfg: Dict[str, Any], batch_size: int) -> Dict[str, T]:
    """The five ``util.sample_uniform(lo, hi, n=batch_size)`` draws of data_modules.py:421-445, in that order."""
    return {k: util.sample_uniform(flanger_cfg[k]["min"], flanger_cfg[k]["max"], n=batch_size)
            for k in FLANGER_PARAM_ORDER}


class FlangerRenderStep:
    """GPU equivalent of ``FlangerCPUDataModule`` 's ``setup`` + ``on_before_batch_transfer``
    (data_modules.py:379-385, 419-458): ``(dry, mod_sig, fx_params)`` in, ``(dry, wet, mod_sig, fx_params)`` out
    with the same fx_params keys.  The x100 upsample of data_modules.py:454-455 is fused into the effect
    kernel; the returned ``mod_sig`` is nevertheless the audio-rate one, as in the reference (data_modules.py:454-458),
    unless ``return_audio_rate_mod=False`` (a consumer that resamples it to 345 frames anyway, lightning.py:114, can
    skip the (B, n_samples) tensor)."""

    def __init__(self, fx_config: Dict[str, Any], batch_size: int, n_samples: int, sr: float,
                 return_audio_rate_mod: bool = True) -> None:
        self.fx_config = fx_config
        self.batch_size = batch_size
        fl = fx_config["flanger"]
        self.flanger = MonoFlangerChorusModule(batch_size=batch_size, n_ch=1, n_samples=n_samples, sr=sr,
                                               max_min_delay_ms=fl["max_min_delay_ms"],
                                               max_lfo_delay_ms=fl["max_lfo_delay_ms"])
        self.return_audio_rate_mod = return_audio_rate_mod

    def __call__(self, batch):
        dry, mod_sig, fx_params = batch
        p = sample_flanger_params(self.fx_config["flanger"], self.batch_size)
        fx_params = dict(fx_params)
        fx_params["depth"] = p["depth"]
        fx_params["feedback"] = p["feedback"]
        fx_params["max_lfo_delay_ms"] = self.flanger.max_lfo_delay_ms
        fx_params["max_min_delay_ms"] = self.flanger.max_min_delay_ms
        fx_params["min_delay_width"] = p["min_delay_width"]
        fx_params["mix"] = p["mix"]
        fx_params["width"] = p["width"]
        if mod_sig.size(-1) == dry.size(-1):
            wet = self.flanger(dry, mod_sig, p["feedback"], p["min_delay_width"], p["width"], p["depth"], p["mix"])
        else:
            wet = self.flanger.forward_control_rate(dry, mod_sig, p["feedback"], p["min_delay_width"], p["width"],
                                                    p["depth"], p["mix"])
            if self.return_audio_rate_mod:
                mod_sig = util.linear_interpolate_last_dim(mod_sig, dry.size(-1))
        return dry, wet, mod_sig, fx_params


# --------------------------------------------------------------------------------------------- N2

def fx_params_md5(f: Dict[str, Any]) -> str:
    """scripts/scratch.py:157-158: md5 of the json of stringified parameter values."""
    hash_dict = {k: str(v.numpy()) if isinstance(v, T) else str(v) for k, v in f.items()}
    return hashlib.md5(json.dumps(hash_dict, sort_keys=True).encode("utf-8")).hexdigest()


def _write_wav_f32(path: str, x: T, sr: int) -> None:
    from scipy.io import wavfile
    a = x.detach().cpu().float().numpy()
    wavfile.write(path, sr, a.T if a.ndim == 2 else a)          # (n_samples, n_ch) 32-bit float WAV


def _read_wav(path: str) -> Tuple[T, int]:
    from scipy.io import wavfile
    sr, a = wavfile.read(path)
    if a.dtype == np.int16:
        a = a.astype(np.float32) / 32768.0
    a = np.atleast_2d(a.astype(np.float32).T if a.ndim == 2 else a.astype(np.float32))
    return tr.from_numpy(np.ascontiguousarray(a)), sr


def write_cache(save_dir: str, dry: T, wet: T, mod_sig: T, fx_params: Dict[str, Any], sr: float) -> List[str]:
    """scripts/scratch.py:145-162: one ``<md5>.pt`` (``{"mod_sig": (n//100,), "fx_params": {...}}``) +
    ``<md5>_dry.wav`` + ``<md5>_wet.wav`` per example.  ``mod_sig`` longer than n//100 is down-sampled like
    scratch.py:150.  Returns the md5 stems."""
    os.makedirs(save_dir, exist_ok=True)
    stems = []
    dry, wet, mod_sig = dry.detach().cpu(), wet.detach().cpu(), mod_sig.detach()
    n_lo = dry.size(-1) // 100
    if mod_sig.size(-1) != n_lo:
        mod_sig = util.linear_interpolate_last_dim(mod_sig, n_lo, align_corners=True)
    mod_sig = mod_sig.cpu()
    for idx in range(dry.size(0)):
        f = {k: v if isinstance(v, float) else v[idx] for k, v in fx_params.items()}
        f = {k: v.item() if isinstance(v, T) else v for k, v in f.items()}
        stem = fx_params_md5(f)
        tr.save({"mod_sig": mod_sig[idx].clone(), "fx_params": f}, os.path.join(save_dir, f"{stem}.pt"))
        _write_wav_f32(os.path.join(save_dir, f"{stem}_dry.wav"), dry[idx], int(sr))
        _write_wav_f32(os.path.join(save_dir, f"{stem}_wet.wav"), wet[idx], int(sr))
        stems.append(stem)
    return stems


class PreprocessedDataset:
    """datasets.py:504-535: reads the cache back as (dry, wet, mod_sig, fx_params)."""

    def __init__(self, input_dir: str, n_samples: int, sr: float) -> None:
        self.input_dir, self.n_samples, self.sr = input_dir, n_samples, sr
        self.pt_paths = sorted(os.path.join(input_dir, f) for f in os.listdir(input_dir) if f.endswith(".pt"))
        self.dry_paths = [f"{p[:-3]}_dry.wav" for p in self.pt_paths]
        self.wet_paths = [f"{p[:-3]}_wet.wav" for p in self.pt_paths]

    def __len__(self) -> int:
        return len(self.pt_paths)

    def __getitem__(self, idx: int):
        data = tr.load(self.pt_paths[idx])
        dry, sr = _read_wav(self.dry_paths[idx])
        assert sr == self.sr
        assert dry.size(-1) == self.n_samples
        wet, sr = _read_wav(self.wet_paths[idx])
        assert sr == self.sr
        assert wet.size(-1) == self.n_samples
        return dry, wet, data["mod_sig"], data["fx_params"]
