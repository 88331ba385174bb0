"""Device-resident timings of the other BASELINE.json configurations (the contract bench.py measures config 4):
  config 1  flanger, one 2 s clip, 2 Hz triangle LFO (the case the reference runs on a CPU in 9.3 s)
  config 2  phaser, 256 x 88 200, random-phase / rate cosine LFO (the ground-truth LFO of datasets.py:442 included)
  config 3  chorus + flanger, 1024 x 88 200, quasi-periodic and distorted control-rate LFOs (LFO generation timed apart)
  config 5  60 s clips x 512: flanger, chorus, phaser (+ the log-mel of one 60 s batch)
CUDA events, 3 warm-ups, median of `reps`.  python scripts/config_bench.py [2] [3] [5]"""
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod_extraction_b200 import modulations as M                      # noqa: E402
from mod_extraction_b200.fx import MonoFlangerChorusModule            # noqa: E402
from mod_extraction_b200.models import LogMelSpectrogram              # noqa: E402
from mod_extraction_b200.phaser import Phaser                         # noqa: E402

dev = torch.device("cuda", 0)
SR = 44100
PEAK = 6547.8
which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 5]
SHAPES6 = ["cos", "tri", "rect_cos", "inv_rect_cos", "saw", "rsaw"]


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, B, N, ms, bytes_per_sample):
    gbs = B * N * bytes_per_sample / (ms * 1e-3) / 1e9
    print(f"{name:46s} {ms:9.3f} ms  {B * N / SR / (ms * 1e-3) / 1e6:7.3f} M audio-s/s  {gbs:7.1f} GB/s algorithmic "
          f"({gbs / PEAK * 100:4.1f} % of {PEAK:.0f})")


def white(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    out = torch.empty((B, 1, N), device=dev)
    step = max(1, (1 << 28) // N)
    for lo in range(0, B, step):
        hi = min(B, lo + step)
        out[lo:hi] = ((torch.rand((hi - lo, 1, N), generator=g) * 2 - 1) * 0.5).to(dev)
    return out


rng = np.random.RandomState(43)
U = lambda B, lo, hi: torch.from_numpy(rng.uniform(lo, hi, B).astype(np.float32)).to(dev)
LU = lambda B, lo, hi: torch.from_numpy(np.exp(rng.uniform(math.log(lo), math.log(hi), B)).astype(np.float32)).to(dev)

if 1 in which:
    # config 1: the reference's own CPU-runnable case (9.3 s in fx.py's python loop): one 2 s clip, flanger, 2 Hz triangle
    B, N = 1, 88200
    x = white(B, N, 42)
    fl1 = MonoFlangerChorusModule(B, 1, N, SR, 1.0, 10.0, check_ranges=False)
    lo = M.make_mod_signal_batch(882, 441.0, [2.0], [0.0], ["tri"])
    out1 = torch.empty_like(x)
    report("config 1: flanger 1 x 2 s, control-rate LFO in", B, N, timed(lambda: fl1.forward_control_rate(x, lo, 0.5, 1.0, 1.0, 1.0, 1.0, out=out1)), 8)
    x_cpu, lo_cpu = x.cpu(), lo.cpu()
    t0 = time.perf_counter()
    for _ in range(5):
        y_cpu = fl1.forward_control_rate(x_cpu, lo_cpu, 0.5, 1.0, 1.0, 1.0, 1.0)
    print(f"          the same through the CPU-tensor drop-in call (H2D + kernel + D2H + sync): {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms")

if 2 in which:
    B, N = 256, 88200
    x = white(B, N, 43)
    rate, depth, fc, fb, mix = LU(B, 0.5, 3), U(B, 0.2, 1), LU(B, 70, 18000), U(B, 0, 0.7), U(B, 0.2, 1)
    ph = Phaser(SR)
    out = torch.empty((B, N), device=dev)
    ms = timed(lambda: ph(x.view(B, N), rate, depth, fc, fb, mix, out=out))
    report("config 2: phaser 256 x 2 s", B, N, ms, 8)
    ms2 = timed(lambda: M.make_mod_signal_batch(N, float(SR), rate.cpu(), np.full(B, math.pi / 2), ["cos"] * B))
    print(f"          + ground-truth cosine LFO at audio rate (datasets.py:442): {ms2:.3f} ms")

if 3 in which:
    B, N = 1024, 88200
    x = white(B, N, 44)
    half = B // 2
    t0 = time.perf_counter()
    base = M.make_mod_signal_batch(882, 441.0, LU(half, 0.5, 2).cpu(), U(half, 0, 2 * math.pi).cpu(), [SHAPES6[i % 6] for i in range(half)])
    torch.manual_seed(44)
    quasi = M.make_quasi_periodic_batch(base, 0.10, 0.3333, 0.10, 0.3333, 0.5)
    dist = M.make_mod_signal_batch(882, 441.0, LU(half, 0.5, 3).cpu(), U(half, 0, 2 * math.pi).cpu(), [SHAPES6[i % 6] for i in range(half)],
                                   np.full(half, 2.0))
    torch.cuda.synchronize()
    lfo_s = time.perf_counter() - t0
    mod_lo = torch.cat([quasi, dist], 0)
    p = [U(B, 0, 0.7), None, U(B, 0.25, 1), U(B, 0.25, 1), U(B, 0.25, 1)]
    out = torch.empty_like(x)
    idx_ch = torch.arange(0, B, 2, dtype=torch.int32, device=dev)        # chorus = even examples, flanger = odd
    idx_fl = torch.arange(1, B, 2, dtype=torch.int32, device=dev)
    ch = MonoFlangerChorusModule(B, 1, N, SR, 30.0, 10.0, check_ranges=False)
    fl = MonoFlangerChorusModule(B, 1, N, SR, 1.0, 10.0, check_ranges=False)
    p_ch = list(p); p_ch[1] = U(B, 0.367, 1)
    p_fl = list(p); p_fl[1] = U(B, 0.0, 1)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for s, m, pp, idx in ((s1, ch, p_ch, idx_ch), (s2, fl, p_fl, idx_fl)):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                m.forward_control_rate(x, mod_lo, *pp, example_index=idx, out=out)
                e = torch.cuda.Event(); e.record(s)
            cur.wait_event(e)

    report("config 3: chorus + flanger 1024 x 2 s", B, N, timed(both), 8)
    report("          chorus half alone", half, N, timed(lambda: ch.forward_control_rate(x, mod_lo, *p_ch, example_index=idx_ch, out=out)), 8)
    report("          flanger half alone", half, N, timed(lambda: fl.forward_control_rate(x, mod_lo, *p_fl, example_index=idx_fl, out=out)), 8)
    print(f"          quasi-periodic + distorted LFO generation (host RNG replay + kernels), once: {lfo_s * 1e3:.1f} ms")

if 5 in which:
    B, N = 512, 2646000
    x = white(B, N, 45)
    out = torch.empty_like(x)
    n_lo = N // 100
    mod_lo = M.make_mod_signal_batch(n_lo, 441.0, LU(B, 0.5, 3).cpu(), U(B, 0, 2 * math.pi).cpu(), [SHAPES6[i % 6] for i in range(B)])
    p = [U(B, 0, 0.7), U(B, 0.367, 1), U(B, 0.25, 1), U(B, 0.25, 1), U(B, 0.25, 1)]
    ch = MonoFlangerChorusModule(B, 1, N, SR, 30.0, 10.0, check_ranges=False)
    fl = MonoFlangerChorusModule(B, 1, N, SR, 1.0, 10.0, check_ranges=False)
    report("config 5: chorus 512 x 60 s", B, N, timed(lambda: ch.forward_control_rate(x, mod_lo, *p, out=out), reps=3), 8)
    p[1] = U(B, 0.0, 1)
    report("config 5: flanger 512 x 60 s", B, N, timed(lambda: fl.forward_control_rate(x, mod_lo, *p, out=out), reps=3), 8)
    ph = Phaser(SR)
    rate, depth, fc, fb, mix = LU(B, 0.5, 3), U(B, 0.2, 1), LU(B, 70, 18000), U(B, 0, 0.7), U(B, 0.2, 1)
    report("config 5: phaser 512 x 60 s", B, N, timed(lambda: ph(x.view(B, N), rate, depth, fc, fb, mix, out=out.view(B, N)), reps=3), 8)
    front = LogMelSpectrogram(SR).to(dev)
    Bm = 64
    lm = torch.empty((Bm, 1, 256, N // 256 + 1), device=dev)
    report("          log-mel of 64 x 60 s rows", Bm, N, timed(lambda: front(x[:Bm], out=lm), reps=3), 8.01)
