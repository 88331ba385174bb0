"""Quick device-time probe of the flanger/chorus kernel (not the contract bench)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mod_extraction_b200.fx import MonoFlangerChorusModule
from mod_extraction_b200.modulations import make_mod_signal_batch

dev = torch.device("cuda", 0)
SR = 44100
def run(name, B, N, mmd, mld, mdw_lo, mode, reps=10):
    g = torch.Generator(device="cpu").manual_seed(1)
    x = ((torch.rand((B, 1, N), generator=g) * 2 - 1) * 0.5).to(dev)
    rng = np.random.RandomState(0)
    f = np.exp(rng.uniform(np.log(0.5), np.log(3.0), B)); ph = rng.uniform(0, 2*math.pi, B)
    shapes = [["cos","tri","rect_cos","inv_rect_cos","saw","rsaw"][b % 6] for b in range(B)]
    n_lo = N // 100
    lo = make_mod_signal_batch(n_lo, 441.0, f, ph, shapes, np.full(B, 2.0))
    U = lambda lo_, hi: torch.from_numpy(rng.uniform(lo_, hi, B).astype(np.float32)).to(dev)
    p = [U(0, 0.7), U(mdw_lo, 1.0), U(0.25, 1), U(0.25, 1), U(0.25, 1)]
    m = MonoFlangerChorusModule(B, 1, N, SR, mmd, mld, check_ranges=False)
    out = torch.empty_like(x)
    if mode == "audio":
        from mod_extraction_b200.util import linear_interpolate_last_dim
        mod = linear_interpolate_last_dim(lo, N)
        fn = lambda: m._render(x, __import__("mod_extraction_b200._ops", fromlist=["ModSource"]).ModSource.audio_rate(mod), *p, None, out)
        bytes_per = 12
    else:
        fn = lambda: m.forward_control_rate(x, lo, *p, out=out)
        bytes_per = 8
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = np.median(ts)
    gbs = B * N * bytes_per / (t * 1e-3) / 1e9
    print(f"{name:28s} B={B:5d} N={N:8d} mode={mode:7s} {t:8.3f} ms  {B*N/SR/(t*1e-3)/1e6:7.3f} M audio-s/s  {gbs:7.1f} GB/s ({gbs/6545*100:5.1f}% of 6545)")

run("chorus", 1024, 88200, 30.0, 10.0, 0.367, "control")
run("flanger", 1024, 88200, 1.0, 10.0, 0.0, "control")
run("flanger audio-rate mod", 1024, 88200, 1.0, 10.0, 0.0, "audio")
run("chorus", 4096, 88200, 30.0, 10.0, 0.367, "control")
run("flanger", 4096, 88200, 1.0, 10.0, 0.0, "control")
run("chorus 60s", 128, 2646000, 30.0, 10.0, 0.367, "control", reps=3)
run("flanger 60s", 128, 2646000, 1.0, 10.0, 0.0, "control", reps=3)
