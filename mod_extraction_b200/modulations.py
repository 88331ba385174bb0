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
           "lfo_kernel_params", "find_corners", "make_quasi_periodic", "make_quasi_periodic_batch",
           "make_combined_mod_sig", "make_combined_mod_sig_batch", "smoothen", "stretch_corners",
           "find_valid_mod_sig_indices", "mod_sig_to_corners"]


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


# --------------------------------------------------------------------------- corners, quasi-periodic, combined

def find_corners(mod_sig: T) -> Tuple[T, T]:
    """modulations.py:219-238.  (B, n) -> (top, bottom), float tensors of 0/1 like the reference."""
    assert mod_sig.ndim == 2
    on_cpu = not mod_sig.is_cuda
    m = mod_sig.detach().float().to(_device()) if on_cpu else mod_sig.detach().float()
    top, bottom = _ops.find_corners(m)
    top, bottom = top.to(mod_sig.dtype), bottom.to(mod_sig.dtype)
    return (top.cpu(), bottom.cpu()) if on_cpu else (top, bottom)


def _stretched_len(size: int, l_min: float, l_max: float, r_min: float, r_max: float, lr_split: float) -> int:
    """Length decision of _time_stretch_section, modulations.py:104-116 (2 host RNG draws)."""
    if util.sample_uniform(0.0, 1.0) < lr_split:
        x = int((util.sample_uniform(l_min, l_max) * size) + 0.5)
        return max(2, size - x)
    x = int((util.sample_uniform(r_min, r_max) * size) + 0.5)
    return size + x


def make_quasi_periodic_batch(mod_sigs: T, l_min: float = 0.2, l_max: float = 0.2, r_min: float = 0.2,
                              r_max: float = 0.2, lr_split: float = 0.5) -> T:
    """make_quasi_periodic for a (B, n) batch: one corner kernel, one host pass that makes the
    reference's RNG draws example by example in the reference's order, one resampling kernel.
    Equivalent to calling the reference function on row 0, then row 1, ... under the same seed."""
    assert mod_sigs.ndim == 2
    dev = mod_sigs.device if mod_sigs.is_cuda else _device()
    m = mod_sigs.detach().float().to(dev)
    B, n = m.shape
    top, bottom = _ops.find_corners(m)
    top, bottom = top.cpu().numpy(), bottom.cpu().numpy()      # the one device->host sync
    sec_off, in_start, in_len, new_len, out_start = [0], [], [], [], []
    for b in range(B):
        corners = top[b] if top[b].sum() > bottom[b].sum() else bottom[b]       # modulations.py:129-132
        idxs = corners.nonzero()[0].tolist()
        if len(idxs) >= 2:                                                       # modulations.py:136-137
            prev, pos = 0, 0
            for idx in idxs:                                                     # modulations.py:142-148
                size = idx + 1 - prev
                nl = _stretched_len(size, l_min, l_max, r_min, r_max, lr_split)
                in_start.append(prev); in_len.append(size); new_len.append(nl); out_start.append(pos)
                pos += nl - 1                                                    # new_section[:-1]
                prev = idx
            tail = n - prev                                                      # modulations.py:150-156
            tail_new = tail + (n - (pos + tail)) if pos + tail < n else tail
            in_start.append(prev); in_len.append(tail); new_len.append(tail_new); out_start.append(pos)
        sec_off.append(len(in_start))
    if not in_start:
        return m.clone()
    return _ops.stretch_sections(m, sec_off, in_start, in_len, new_len, out_start)


def make_quasi_periodic(mod_sig: T,
                        l_min: float = 0.2,
                        l_max: float = 0.2,
                        r_min: float = 0.2,
                        r_max: float = 0.2,
                        lr_split: float = 0.5) -> T:
    """modulations.py:121-160."""
    assert mod_sig.ndim == 1
    on_cpu = not mod_sig.is_cuda
    out = make_quasi_periodic_batch(mod_sig.unsqueeze(0), l_min, l_max, r_min, r_max, lr_split)[0]
    return out.cpu() if on_cpu else out


_PINNED_PAIRS: list = []        # recycled pinned (2,) int32 buffers of the deferred read-back
_PINNED_STAGES: list = []       # (pinned float32 staging buffer, event after its last copy to the device)


def _event_now():
    ev = tr.cuda.Event()
    ev.record()
    return ev


def _pinned_stage(n: int) -> T:
    """A pinned float32 buffer of n elements whose previous host-to-device copy has completed."""
    for i, (buf, ev) in enumerate(_PINNED_STAGES):
        if buf.numel() == n and ev.query():
            del _PINNED_STAGES[i]
            return buf
    if len(_PINNED_STAGES) > 8:
        del _PINNED_STAGES[0]
    return tr.empty((n,), dtype=tr.float32).pin_memory()


def make_combined_mod_sig_batch(n_samples: int, sr: float, freqs, phases, shapes: List[str], device=None,
                                return_base: bool = False, host_replay: bool = False, deferred: bool = False):
    """make_combined_mod_sig for B (freq, phase) pairs, equivalent to calling the reference function on pair 0, then
    pair 1, ... under the same state of the torch global CPU generator, which it leaves where that loop would.

    The number of draws of an example depends on the corners of the base shape it drew.  The generator's next raw
    words are produced in one go (``_rng.TorchMT``), every candidate base shape is rendered and corner-searched on
    the GPU, one device thread replays the reference's draw order, and the generator is advanced by the number of
    words that replay consumed: one launch sequence and one 8-byte read-back for the whole batch instead of a python
    loop of scalar ``torch.randint`` calls (0.2 s for 4096 examples).  ``host_replay=True`` keeps that loop (the
    cross-check of the device replay, and the fallback for signals with more than 64 bottom corners).

    ``deferred=True`` returns ``(out, finish)`` right after the launches: the caller queues whatever consumes ``out`` and
    then calls ``finish()``, which performs the 8-byte read-back, advances the generator and returns True -- or False
    when the device replay could not be used (more words or corners than provisioned: practically never), in which case
    ``out`` is invalid, the generator is untouched and the caller falls back to the blocking call."""
    device = _device() if device is None else device
    f = tr.as_tensor(freqs, dtype=tr.float64).reshape(-1)
    p = tr.as_tensor(phases, dtype=tr.float64).reshape(-1)
    B, S = f.numel(), len(shapes)
    assert B == p.numel() and S > 0
    assert bool((f > 0.0).all()) and bool((f < sr / 2.0).all())             # modulations.py:23
    assert bool((p >= -2 * tr.pi).all()) and bool((p <= 2 * tr.pi).all())   # modulations.py:24
    sid = shape_ids(shapes)
    if deferred:
        assert not host_replay and not return_base and B > 0 and 3 <= n_samples <= 32767
        from ._rng import TorchMT
        mt = TorchMT()
        # every host operand goes through ONE pinned buffer and one asynchronous copy: a copy from pageable memory
        # would make the host wait for whatever the stream is still running (the previous step's render)
        n_words = B * 17
        stage = _pinned_stage(2 * B + S + n_words)
        stage[:B].copy_(f.float())
        stage[B:2 * B].copy_(p.float())
        stage[2 * B:2 * B + S].view(tr.int32).copy_(sid)
        stage[2 * B + S:].view(tr.int32).copy_(tr.from_numpy(mt.words(n_words).view("int32")))
        on_dev = stage.to(device, non_blocking=True)
        _PINNED_STAGES.append((stage, _event_now()))
        out, _, consumed = _ops.combined_lfo(n_samples, sr, on_dev[:B], on_dev[B:2 * B], on_dev[2 * B:2 * B + S].view(tr.int32),
                                             on_dev[2 * B + S:].view(tr.int32))

        # the read-back is queued right behind the replay (pinned destination + event), so finish() waits for the LFO
        # kernels only, not for whatever the caller queues after them
        back = _PINNED_PAIRS.pop() if _PINNED_PAIRS else tr.empty((2,), dtype=tr.int32).pin_memory()
        back.copy_(consumed, non_blocking=True)
        landed = tr.cuda.Event()
        landed.record()

        def finish() -> bool:
            landed.synchronize()
            used, err = back.tolist()
            _PINNED_PAIRS.append(back)
            if err == 0:
                mt.consume(used)
            return err == 0
        return out, finish
    if not host_replay and B > 0 and 3 <= n_samples <= 32767:
        from ._rng import TorchMT
        mt = TorchMT()
        per_example = 1 + 16                # base draw + sections; a 2 s control-rate LFO below 3 Hz has at most 7
        for _ in range(2):
            words = tr.from_numpy(mt.words(B * per_example).view("int32"))
            out, base, consumed = _ops.combined_lfo(n_samples, sr, f.float().to(device, non_blocking=True),
                                                    p.float().to(device, non_blocking=True), sid.to(device),
                                                    words.to(device, non_blocking=True))
            used, err = consumed.tolist()
            if err == 0:
                mt.consume(used)
                return (out, base) if return_base else out
            if err != 1:
                break
            mt = TorchMT()                  # ran out of words: hand over more and replay
            per_example *= 8
    cand = make_mod_signal_batch(n_samples, sr, f.repeat_interleave(S), p.repeat_interleave(S), sid.repeat(B),
                                 None, device)                                   # (B*S, n)
    _, bottom = _ops.find_corners(cand)
    bottom = bottom.cpu().numpy().reshape(B, S, n_samples)
    base = []
    sec_off, sec_start, sec_len, sec_shape = [0], [], [], []
    for b in range(B):
        k = util.randint(0, S)                                                   # util.choice, modulations.py:196
        base.append(b * S + k)
        idxs = bottom[b, k].nonzero()[0].tolist()
        if len(idxs) > 1:                                                        # modulations.py:203-209
            for i, idx in enumerate(idxs[1:]):
                prev = idxs[i]
                section_len = idx - prev + 1
                shape = shapes[util.randint(0, S)]
                assert 0.0 < 1.0 < section_len / 2.0                             # make_mod_signal's assert :23
                sec_start.append(prev); sec_len.append(section_len); sec_shape.append(SHAPE_ID[shape])
        sec_off.append(len(sec_start))
    out = cand[tr.tensor(base, device=device)].contiguous()
    if sec_start:
        _ops.lfo_sections_(out, sec_off, sec_start, sec_len, sec_shape)
    if return_base:
        return out, (tr.tensor(base, dtype=tr.int32) % S).to(device)
    return out


def make_combined_mod_sig(n_samples: int,
                          sr: float,
                          freq: float,
                          phase: float,
                          shapes: List[str],
                          device=None) -> T:
    """modulations.py:191-210."""
    return make_combined_mod_sig_batch(n_samples, sr, [float(freq)], [float(phase)], shapes, device)[0]


# --------------------------------------------------------------------------- extracted-LFO post-processing
# (eval path of the reference: lightning.py:114-127,284-300,325-337)

def _to_dev(x: T) -> Tuple[T, bool]:
    on_cpu = not x.is_cuda
    return (x.detach().float().to(_device()) if on_cpu else x.detach().float()), on_cpu


def smoothen(x: T, smooth_n_frames: int) -> T:
    """modulations.py:358-362: moving average over the last dim, (..., n) -> (..., n - smooth_n_frames + 1)."""
    if smooth_n_frames <= 1:
        return x
    xd, on_cpu = _to_dev(x)
    lead = xd.shape[:-1]
    out = _ops.smoothen(xd.reshape(-1, xd.size(-1)), smooth_n_frames).reshape(lead + (-1,))
    return out.cpu() if on_cpu else out


def mod_sig_to_corners(mod_sig: T, n_frames: int) -> Tuple[T, T]:
    """modulations.py:212-215."""
    assert mod_sig.ndim == 2
    return find_corners(util.linear_interpolate_last_dim(mod_sig, n_frames, align_corners=True))


def stretch_corners(mod_sig: T, max_n_corners: int = 10, smooth_n_frames: int = 32) -> T:
    """modulations.py:294-307: smooth, find the corners and stretch every row so that its top corners
    reach 1 and its bottom corners 0; rows with more than ``max_n_corners`` corners are only smoothed.
    Two launches for the whole batch instead of a python loop over rows and corners."""
    assert mod_sig.ndim == 2
    xd, on_cpu = _to_dev(mod_sig)
    if smooth_n_frames > 1:
        xd = _ops.smoothen(xd, smooth_n_frames)
    out = _ops.stretch_corners(xd, max_n_corners)
    return out.cpu() if on_cpu else out


def find_valid_mod_sig_indices(mod_sig: T, min_top_corners: int = 1, max_top_corners: int = 6,
                               min_bottom_corners: int = 1, max_bottom_corners: int = 6,
                               min_fraction_between_corners: float = 0.10) -> List[int]:
    """modulations.py:348-355 with the thresholds of check_mod_sig (:311-345) as keywords."""
    assert mod_sig.ndim == 2
    xd, _ = _to_dev(mod_sig)
    min_n_frames = int(min_fraction_between_corners * xd.size(-1))
    valid = _ops.check_mod_sig(xd, min_top_corners, max_top_corners, min_bottom_corners, max_bottom_corners,
                               min_n_frames)
    return tr.nonzero(valid, as_tuple=True)[0].tolist()
