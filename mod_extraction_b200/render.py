"""Batch renderer for the ``train_lfo_interwoven_all`` data path (BASELINE config 4).

Reference: ``InterwovenDataset.__getitem__`` picks ``datasets[idx % 3]`` -- flanger, chorus, phaser in
the order of configs/data/interwoven_idmt_all.yml:12-17 (datasets.py:79-83); the flanger / chorus
examples are rendered by ``MonoFlangerChorusModule`` from a control-rate LFO (data_modules.py:419-458),
the phaser examples by pedalboard (datasets.py:455-482), and ``Spectral2DCNN.forward`` takes the log-mel
of ``cat([dry, wet], dim=1)`` (lightning.py:106, models.py:199-208).

Here one call renders the whole interleaved batch on the GPU: three chains "effect -> log-mel of its wet rows" on
three high-priority streams (examples are selected through index lists, nothing is gathered or copied) and the
log-mel of the dry half, which depends on nothing, on a low-priority fourth stream that fills the gaps -- all
writing straight into the (B, 2, n_mels, n_frames) tensor the extractor consumes.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _ops
from ._ops import ModSource
from .fx import _delay_samples
from .models import LogMelSpectrogram

FLANGER, CHORUS, PHASER = 0, 1, 2       # order of configs/data/interwoven_idmt_all.yml


class _HostStep:
    """One queued host-buffer step (InterwovenRenderer.render_host(wait=False))."""

    def __init__(self, end_events, keep, stream) -> None:
        self._end, self._keep, self._stream = end_events, keep, stream

    def wait(self) -> None:
        """Blocks until every output of the step has landed in its host buffer; later work on the caller's stream is
        ordered behind the step as well."""
        if self._end is None:
            return
        for ev in self._end:
            self._stream.wait_event(ev)
            ev.synchronize()
        self._end, self._keep = None, None                  # device staging buffers may go back to the allocator


class InterwovenRenderer:
    def __init__(self, n_samples: int = 88200, sr: float = 44100.0, device=None,
                 flanger_delays_ms: Tuple[float, float] = (1.0, 10.0),      # configs/data/gen_idmt_fl.yml
                 chorus_delays_ms: Tuple[float, float] = (30.0, 10.0),      # configs/data/gen_idmt_ch.yml
                 n_mels: int = 256, n_fft: int = 1024, hop_len: int = 256, eps: float = 1e-7,
                 phaser_buffer_size: int = 8192, concurrent: bool = True) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("mod_extraction_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_samples, self.sr = n_samples, sr
        self.fl = tuple(_delay_samples(ms, sr) for ms in flanger_delays_ms)
        self.ch = tuple(_delay_samples(ms, sr) for ms in chorus_delays_ms)
        self.front = LogMelSpectrogram(int(sr), n_fft, hop_len, n_mels, eps=eps).to(self.device)
        self.n_frames = self.front.n_frames(n_samples)
        self.phaser_buffer_size = phaser_buffer_size
        self.concurrent = concurrent
        # three effect chains (effect -> log-mel of its wet rows) on high-priority streams, the dry log-mel (which
        # depends on nothing) on a low-priority one: the hardware fills the gaps of the chains with it
        self._streams = ([torch.cuda.Stream(device=self.device, priority=-1) for _ in range(3)] +
                         [torch.cuda.Stream(device=self.device, priority=0)]) if concurrent else None
        self._idx_cache: Dict[Tuple[int, bytes], Tuple[Tensor, ...]] = {}

    # ------------------------------------------------------------------ helpers
    def _groups(self, effect: Tensor, lo: int = 0, hi: Optional[int] = None):
        """Index lists (relative to `lo`) of the flanger / chorus / phaser examples in effect[lo:hi].
        Cached by CONTENT (the bytes of effect[lo:hi]), so a caller that refills one `effect` buffer in place
        between steps gets fresh lists; values outside {0, 1, 2} are rejected (their rows would stay unwritten)."""
        hi = effect.numel() if hi is None else hi
        e = effect.detach().reshape(-1)[lo:hi].to("cpu", torch.int64).contiguous()
        key = (hi - lo, e.numpy().tobytes())
        hit = self._idx_cache.get(key)
        if hit is not None:
            return hit
        if e.numel() and (int(e.min()) < FLANGER or int(e.max()) > PHASER):
            raise ValueError("effect ids must be 0 (flanger), 1 (chorus) or 2 (phaser)")
        groups = []
        for k in (FLANGER, CHORUS, PHASER):
            idx = torch.nonzero(e == k).reshape(-1).to(torch.int32)
            groups.append(idx.to(self.device))
        not_phaser = torch.nonzero(e != PHASER).reshape(-1).to(self.device)        # int64: rows whose dry audio is given
        if len(self._idx_cache) > 64:
            self._idx_cache.clear()
        self._idx_cache[key] = (*groups, not_phaser)
        return self._idx_cache[key]

    def alloc_outputs(self, B: int) -> Tuple[Tensor, Tensor]:
        wet = torch.empty((B, 1, self.n_samples), device=self.device, dtype=torch.float32)
        logmel = torch.empty((B, 2, self.front.n_mels, self.n_frames), device=self.device, dtype=torch.float32)
        return wet, logmel

    # ------------------------------------------------------------------ the hot path
    @staticmethod
    def _chunk_edges(B: int, chunk: int):
        """Chunk schedule of the host-buffer path: a short first and last chunk keep the part of the pipeline that cannot
        overlap (the first H2D copy, the last kernels + D2H copy) small; everything in between moves `chunk` examples."""
        edges, lo = [], 0
        short = max(1, chunk // 4)
        while lo < B:
            left = B - lo
            size = short if (lo == 0 and B > 2 * chunk) else (left if left <= chunk + short else chunk)
            if lo > 0 and left <= chunk + short and left > short and B > 2 * chunk:
                edges.append((lo, lo + left - short))
                lo += left - short
                size = short
            edges.append((lo, lo + size))
            lo += size
        return edges

    @torch.no_grad()
    def render_host(self, dry_h: Tensor, effect: Tensor, mod_lo_h, fc_h: Dict[str, Tensor],
                    ph_h: Dict[str, Tensor], wet_h: Tensor, logmel: Tensor, stat_h: Optional[Tensor] = None,
                    chunk: int = 512, dry_d: Optional[Tensor] = None, wet_d: Optional[Tensor] = None,
                    logmel_h: Optional[Tensor] = None, ph_long_h: Optional[Tensor] = None,
                    ph_start_h: Optional[Tensor] = None, dry_ph_h: Optional[Tensor] = None,
                    dry_fc_h: Optional[Tensor] = None, ph_packed_h: Optional[Tensor] = None,
                    ph_offsets=None, wait: bool = True, duplex: bool = True):
        """Host-buffer entry point: pinned host dry audio + parameters in, wet audio out to the pinned host
        tensor `wet_h`; the log-mel tensor stays on the GPU (`logmel`, (B,2,n_mels,n_frames)) where the
        extractor consumes it, and `stat_h` (B, 2) receives its per-example mean; with `logmel_h` (pinned, same
        shape) the log-mel tensor is delivered to the host as well.

        `mod_lo_h`: the (B, n_lo) control-rate LFOs -- a pinned host tensor, a device tensor, or a CALLABLE returning the
        device tensor (LFO synthesis on the device): the callable runs after every input copy has been queued, so the
        synthesis hides behind the first copies instead of delaying them.  The callable may also return
        ``(tensor, finish)`` (``make_combined_mod_sig_batch(..., deferred=True)``): ``finish()`` -- the read-back that
        advances the host generator -- is then called after all the work of the step has been queued; if it returns
        False the whole step is redone with ``mod_lo_h(blocking=True)``.
        Phaser examples (rendered over a longer chunk and cropped, see `render`) come either as `ph_long_h`
        (n_phaser, L) pinned rows of a common pitch, or -- what a collate function produces from the variable-length
        chunks the reference's dataset reads -- as `ph_packed_h`, a 1-D pinned array of the rows back to back, with
        `ph_offsets` (n_phaser + 1 host ints, multiples of 4) their starts; row i then only needs its first
        start + N samples, the causal prefix that determines the window.  `ph_start_h` (B,) int32: the window starts.
        The dry windows cut out of the phaser chunks are written into the phaser rows of `dry_d` on the device (the
        log-mel needs them); they are views of host data the caller already has, so they are only copied back when
        `dry_ph_h` (n_phaser, N) pinned is given.  `dry_fc_h` (n_flanger + n_chorus, N) pinned: the dry audio of the other
        examples as a compact array in batch order -- the per-effect layout the reference's three datasets produce --
        in which case `dry_h` is not read at all (pass a (B, 1, N) meta / empty tensor for the shape) and the rows are
        scattered into the interleaved batch on the device.

        Schedule: the input copies are queued ahead of everything else on the copy-in stream (one event per chunk), then
        per chunk the kernels (compute stream) and the output copies (copy-out stream), so the H2D engine never waits for
        the host and the H2D copy of chunk i+k, the kernels of chunk i and the D2H copy of chunk i-1 overlap (PCIe is
        full duplex); the call returns when everything has landed.

        ``wait=False`` returns a handle right after queuing instead (``handle.wait()`` blocks until this step's outputs
        have landed): the next call may be issued before that, and its input copies then run while this step's last
        output copies drain -- consecutive steps are pipelined like chunks are.  The device buffers ``dry_d`` / ``wet_d``
        / ``logmel`` may be the same in consecutive calls (per-chunk events order the reuse); the HOST output buffers
        must not be read before ``handle.wait()`` and should alternate between two sets when steps overlap.

        ``duplex=False`` holds every output copy back until the last input copy of the step has finished: the link then
        carries one direction at a time.  Slower on a host that sustains full duplex (one rank: both directions overlap
        almost completely), faster where several ranks share a host whose copy rate collapses when both directions run
        (measured with eight ranks: 233 GB/s in alone, 120 GB/s out alone, 77 GB/s per direction together)."""
        B, _, N = dry_h.shape
        dev = self.device
        if B == 0:
            return None if wait else _HostStep([], [], torch.cuda.current_stream(dev))
        if dry_d is None:
            dry_d = torch.empty((B, 1, N), device=dev, dtype=torch.float32)
        if wet_d is None:
            wet_d = torch.empty((B, 1, N), device=dev, dtype=torch.float32)
        if not hasattr(self, "_io_streams"):
            self._io_streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
        s_in, s_run, s_out = self._io_streams
        cur = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(cur)                                   # whatever the caller queued before this call
        for st in self._io_streams:
            st.wait_event(start)
        prev = getattr(self, "_prev_host_step", None)       # a step that may still be in flight (wait=False)
        fc_keys = ("feedback", "min_delay_width", "width", "depth", "mix")
        ph_keys = ("rate_hz", "depth", "centre_frequency_hz", "feedback", "mix")
        edges = self._chunk_edges(B, chunk)
        packed = ph_packed_h is not None
        cropped = packed or ph_long_h is not None
        assert not (packed and ph_long_h is not None), "phaser chunks come packed OR padded"
        if cropped:
            assert ph_start_h is not None
            is_ph = (effect.detach().reshape(-1).cpu() == PHASER)
            ph_before = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(is_ph.to(torch.int64), 0)]).tolist()
        if packed:
            offs = [int(v) for v in ph_offsets]
            assert len(offs) == ph_before[B] + 1 and all(o % 4 == 0 for o in offs), "ph_offsets: n_phaser + 1 multiples of 4"
            max_len = int(ph_start_h.max()) + N if ph_start_h.numel() else N
        assert dry_fc_h is None or cropped, "dry_fc_h goes with the phaser chunks"
        keep = []                                           # device buffers that must outlive the queued work

        # ---- input copies of one chunk on the copy-in stream
        with torch.cuda.stream(s_in):
            f_all = {k: fc_h[k].to(dev, non_blocking=True) for k in fc_keys}
            p_all = {k: ph_h[k].to(dev, non_blocking=True) for k in ph_keys}
            pst_all = ph_start_h.to(dev, non_blocking=True) if cropped else None
            offs_all = torch.tensor(offs, dtype=torch.int64).pin_memory().to(dev, non_blocking=True) if packed else None
            m_all = None
            if not callable(mod_lo_h):
                m_all = mod_lo_h if mod_lo_h.is_cuda else mod_lo_h.to(dev, non_blocking=True)
        keep += [*f_all.values(), *p_all.values(), pst_all, offs_all]

        same_plan = prev is not None and prev["edges"] == edges
        if prev is not None and not same_plan:              # other chunking: order behind the whole previous step
            for st in self._io_streams:
                for ev in prev["end"]:
                    st.wait_event(ev)
        run_events, out_events = [], []

        def queue_in(lo, hi):
            if same_plan:                                   # dry_d[lo:hi] is free once the previous step's kernels read it
                s_in.wait_event(prev["run"][edges.index((lo, hi))])
            with torch.cuda.stream(s_in):
                if dry_fc_h is None:
                    dry_d[lo:hi].copy_(dry_h[lo:hi], non_blocking=True)
                else:
                    f0, f1 = lo - ph_before[lo], hi - ph_before[hi]
                    stage = dry_fc_h[f0:f1].to(dev, non_blocking=True)
                    dry_d[lo:hi].view(hi - lo, N).index_copy_(0, self._groups(effect, lo, hi)[3], stage)
                    keep.append(stage)
                pl = po = None
                r0 = r1 = 0
                if cropped:
                    r0, r1 = ph_before[lo], ph_before[hi]
                    if packed:
                        pl = ph_packed_h[offs[r0]:offs[r1]].to(dev, non_blocking=True)
                        po = offs_all[r0:r1] - offs[r0]
                    else:
                        pl = ph_long_h[r0:r1].to(dev, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            keep.extend((pl, po))
            return ev_in, pl, po, r0, r1

        # ---- kernels of one chunk on the compute stream, then its output copies on the copy-out stream
        def run(lo, hi, ev_in, pl, po, r0, r1):
            s_run.wait_event(ev_in)
            if same_plan:                                   # wet_d / logmel[lo:hi] are free once their last copies left
                s_run.wait_event(prev["out"][edges.index((lo, hi))])
            with torch.cuda.stream(s_run):
                f = {k: v[lo:hi] for k, v in f_all.items()}
                p = {k: v[lo:hi] for k, v in p_all.items()}
                self.render(dry_d[lo:hi], effect, m_all[lo:hi], f, p, wet=wet_d[lo:hi], logmel=logmel[lo:hi], _range=(lo, hi),
                            ph_long=pl, ph_start=None if pst_all is None else pst_all[lo:hi], ph_offsets=po,
                            ph_max_len=max_len if packed else None)
                st_d = logmel[lo:hi].mean(dim=(2, 3)) if stat_h is not None else None
                dph = None
                if cropped and dry_ph_h is not None and r1 > r0:
                    dph = dry_d[lo:hi].view(hi - lo, N).index_select(0, self._groups(effect, lo, hi)[2].to(torch.int64))
                ev_run = torch.cuda.Event()
                ev_run.record(s_run)
            s_out.wait_event(ev_run)
            if not duplex:
                s_out.wait_event(staged[-1][0])             # the last input copy of the step
            with torch.cuda.stream(s_out):
                wet_h[lo:hi].copy_(wet_d[lo:hi], non_blocking=True)
                if logmel_h is not None:
                    logmel_h[lo:hi].copy_(logmel[lo:hi], non_blocking=True)
                if stat_h is not None:
                    stat_h[lo:hi].copy_(st_d, non_blocking=True)
                    keep.append(st_d)
                if dph is not None:
                    dry_ph_h[r0:r1].copy_(dph, non_blocking=True)
                    keep.append(dph)
                ev_out = torch.cuda.Event()
                ev_out.record(s_out)
            run_events.append(ev_run)
            out_events.append(ev_out)

        # ---- schedule: two chunks of input copies go out first (they keep the H2D engine busy for the next few
        # milliseconds), then the LFO synthesis and the first chunk's kernels -- so the first output copy starts as early
        # as it can -- then every remaining input copy in one go, then the remaining chunks
        ahead = min(2, len(edges)) if duplex else len(edges)
        staged = [queue_in(*e) for e in edges[:ahead]]
        finish = None
        if callable(mod_lo_h):
            with torch.cuda.stream(s_run):
                m_all = mod_lo_h()                          # device-side LFO synthesis, behind the first input copies
                if isinstance(m_all, tuple):
                    m_all, finish = m_all
        run(*edges[0], *staged[0])
        staged += [queue_in(*e) for e in edges[ahead:]]
        for e, st in zip(edges[1:], staged[1:]):
            run(*e, *st)
        end = []
        for st in self._io_streams:
            ev = torch.cuda.Event()
            ev.record(st)
            end.append(ev)
        # read-back of the LFO synthesis (advances the host generator): its kernels ran long before the host got here
        ok = finish() if finish is not None else True
        self._prev_host_step = {"edges": edges, "run": run_events, "out": out_events, "end": end}
        handle = _HostStep(end, keep, cur)
        if not ok:                                          # device replay not usable (practically never): redo, blocking
            handle.wait()
            self._prev_host_step = None
            m = mod_lo_h(blocking=True)
            return self.render_host(dry_h, effect, m, fc_h, ph_h, wet_h, logmel, stat_h, chunk, dry_d, wet_d, logmel_h,
                                    ph_long_h, ph_start_h, dry_ph_h, dry_fc_h, ph_packed_h, ph_offsets, wait, duplex)
        if wait:
            handle.wait()
            return None
        return handle

    @torch.no_grad()
    def render(self, dry: Tensor, effect: Tensor, mod_lo: Tensor, fc: Dict[str, Tensor], ph: Dict[str, Tensor],
               wet: Optional[Tensor] = None, logmel: Optional[Tensor] = None,
               _range: Optional[Tuple[int, int]] = None, ph_long: Optional[Tensor] = None,
               ph_start: Optional[Tensor] = None, ph_offsets: Optional[Tensor] = None,
               ph_max_len: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        """dry (B,1,N) CUDA float32; effect (B,) ints in {0: flanger, 1: chorus, 2: phaser};
        mod_lo (B, n_lo) control-rate LFO of the flanger / chorus examples (rows of phaser examples are
        ignored); fc: feedback, min_delay_width, width, depth, mix as (B,) tensors (data_modules.py:421-445);
        ph: rate_hz, depth, centre_frequency_hz, feedback, mix as (B,) tensors (datasets.py:461-470).
        Returns (wet (B,1,N), logmel (B,2,n_mels,n_frames)).

        The reference renders a phaser example over ``N + round(sr / rate_hz)`` samples and keeps a random window of N
        (PedalboardPhaserDataset.__getitem__, datasets.py:428-447).  ``ph_long`` (n_phaser, L) holds those longer chunks,
        one row per phaser example in batch order, and ``ph_start`` (B,) the window starts: the phaser then runs over the
        long rows and writes the window of its wet output into ``wet`` AND the same window of the dry chunk into the
        phaser rows of ``dry`` (in place), before their log-mel is taken.  Without ``ph_long`` the phaser rows of ``dry``
        are rendered as they are (no extra period).  With ``ph_offsets`` (int64, one per phaser example) ``ph_long`` is a
        packed 1-D array instead: row i starts at ``ph_long[ph_offsets[i]]`` and holds its first ``start + N`` samples;
        ``ph_max_len`` bounds ``start + N``."""
        assert dry.is_cuda and dry.dtype == torch.float32 and dry.ndim == 3 and dry.size(1) == 1
        assert dry.is_contiguous()
        B, _, N = dry.shape
        assert N == self.n_samples
        if wet is None or logmel is None:
            w2, l2 = self.alloc_outputs(B)
            wet = w2 if wet is None else wet
            logmel = l2 if logmel is None else logmel
        i_fl, i_ch, i_ph, _ = self._groups(effect) if _range is None else self._groups(effect, *_range)
        fc_args = [fc[k] for k in ("feedback", "min_delay_width", "width", "depth", "mix")]
        ph_args = [ph[k] for k in ("rate_hz", "depth", "centre_frequency_hz", "feedback", "mix")]
        src = ModSource.control_rate(mod_lo)
        nm = self.front.n_mels * self.n_frames
        dry2, wet2 = dry.view(B, N), wet.view(B, N)
        lm_dry, lm_wet = logmel.view(-1), logmel.view(-1)[nm:]
        cropped = ph_long is not None and i_ph.numel() > 0
        if cropped:
            assert ph_start is not None and ph_long.is_cuda and ph_start.numel() == B
            assert (ph_long.shape[0] == i_ph.numel()) if ph_offsets is None else (ph_offsets.numel() == i_ph.numel())

        def effects_fl():
            _ops.flanger_chorus(dry, src, self.fl[0], self.fl[1], *fc_args, example_index=i_fl, out=wet)

        def effects_ch():
            _ops.flanger_chorus(dry, src, self.ch[0], self.ch[1], *fc_args, example_index=i_ch, out=wet)

        def effects_ph():
            if cropped:
                _ops.phaser_crop(ph_long, N, ph_start, self.sr, *ph_args, block=self.phaser_buffer_size,
                                 example_index=i_ph, out=wet2, dry_out=dry2, row_offsets=ph_offsets, max_len=ph_max_len)
            else:
                _ops.phaser(dry2, self.sr, *ph_args, block=self.phaser_buffer_size, example_index=i_ph, out=wet2)

        def mel(x2, out_flat, rows):
            if rows is not None and rows.numel() == 0:
                return
            self.front.forward_rows(x2, N, B, out_flat, N, 2 * nm, rows)

        if not self.concurrent:
            effects_fl(); effects_ch(); effects_ph()
            mel(dry2, lm_dry, None)
            mel(wet2, lm_wet, None)
            return wet, logmel

        cur = torch.cuda.current_stream(self.device)
        s_fl, s_ch, s_ph, s_dry = self._streams
        start = torch.cuda.Event()
        start.record(cur)
        for s, fn, rows in ((s_fl, effects_fl, i_fl), (s_ch, effects_ch, i_ch), (s_ph, effects_ph, i_ph)):
            s.wait_event(start)
            with torch.cuda.stream(s):
                fn()
                mel(wet2, lm_wet, rows)                   # same stream: starts when its own effect has finished
                if cropped and rows is i_ph:
                    mel(dry2, lm_dry, rows)               # the dry window of the phaser rows exists only now
                ev = torch.cuda.Event()
                ev.record(s)
            cur.wait_event(ev)
        s_dry.wait_event(start)
        with torch.cuda.stream(s_dry):
            if cropped:                                   # needs no effect: background work for the whole step
                mel(dry2, lm_dry, i_fl)
                mel(dry2, lm_dry, i_ch)
            else:
                mel(dry2, lm_dry, None)
            fin = torch.cuda.Event()
            fin.record(s_dry)
        cur.wait_event(fin)
        for t in (dry, mod_lo, wet, logmel, ph_long, ph_start, ph_offsets, *fc_args, *ph_args):
            if isinstance(t, Tensor) and t.is_cuda:
                for s in self._streams:
                    t.record_stream(s)
        return wet, logmel
