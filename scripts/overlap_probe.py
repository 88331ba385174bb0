"""How well do the four streams of InterwovenRenderer.render overlap?  Times the step serialised on one stream,
concurrent, and its parts alone (config-4 shape).  python scripts/overlap_probe.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                   # noqa: E402
from mod_extraction_b200 import _ops                           # noqa: E402
from mod_extraction_b200._ops import ModSource                 # noqa: E402
from mod_extraction_b200.render import InterwovenRenderer      # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
torch.manual_seed(0)
N = bench.N
dry = (torch.rand(B, 1, N, device=dev) - 0.5)
effect = torch.arange(B) % 3
mod_lo = torch.rand(B, 882, device=dev)
U = lambda lo, hi: (torch.rand(B, device=dev) * (hi - lo) + lo)
fc = {"feedback": U(0, 0.7), "min_delay_width": U(0, 1), "width": U(0.25, 1), "depth": U(0.25, 1), "mix": U(0.25, 1)}
ph = {"rate_hz": U(0.5, 3), "depth": U(0.2, 1), "centre_frequency_hz": U(70, 18000), "feedback": U(0, 0.7), "mix": U(0.2, 1)}


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for conc in (False, True):
    R = InterwovenRenderer(N, float(bench.SR), dev, concurrent=conc)
    wet, logmel = R.alloc_outputs(B)
    ms = timed(lambda: R.render(dry, effect, mod_lo, fc, ph, wet=wet, logmel=logmel))
    print(f"concurrent={conc}: {ms:.3f} ms per step")
R = InterwovenRenderer(N, float(bench.SR), dev, concurrent=True)
wet, logmel = R.alloc_outputs(B)
i_fl, i_ch, i_ph, _ = R._groups(effect)
src = ModSource.control_rate(mod_lo)
fa = [fc[k] for k in ("feedback", "min_delay_width", "width", "depth", "mix")]
pa = [ph[k] for k in ("rate_hz", "depth", "centre_frequency_hz", "feedback", "mix")]
nm = 256 * R.n_frames
parts = {
    "flanger": lambda: _ops.flanger_chorus(dry, src, R.fl[0], R.fl[1], *fa, example_index=i_fl, out=wet),
    "chorus": lambda: _ops.flanger_chorus(dry, src, R.ch[0], R.ch[1], *fa, example_index=i_ch, out=wet),
    "phaser": lambda: _ops.phaser(dry.view(B, N), float(bench.SR), *pa, block=8192, example_index=i_ph, out=wet.view(B, N)),
    "logmel dry (B rows)": lambda: R.front.forward_rows(dry.view(B, N), N, B, logmel.view(-1), N, 2 * nm, None),
    "logmel wet, 3 launches": lambda: [R.front.forward_rows(wet.view(B, N), N, B, logmel.view(-1)[nm:], N, 2 * nm, r)
                                       for r in (i_ch, i_ph, i_fl)],
    "logmel wet, 1 launch": lambda: R.front.forward_rows(wet.view(B, N), N, B, logmel.view(-1)[nm:], N, 2 * nm, None),
}
tot = 0.0
for k, fn in parts.items():
    ms = timed(fn)
    print(f"{k:26s} {ms:.3f} ms")
streams = [torch.cuda.Stream(device=dev) for _ in range(3)]


def effects_concurrent():
    cur = torch.cuda.current_stream()
    ev0 = torch.cuda.Event(); ev0.record(cur)
    for s, k in zip(streams, ("flanger", "chorus", "phaser")):
        s.wait_event(ev0)
        with torch.cuda.stream(s):
            parts[k]()
            ev = torch.cuda.Event(); ev.record(s)
        cur.wait_event(ev)


print(f"three effects on three streams: {timed(effects_concurrent):.3f} ms")
