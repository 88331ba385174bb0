"""Is the 512-example step (config 4 cut over 8 GPUs) bound by the host?  Device time per step (CUDA events over K steps)
against the host time it takes to QUEUE a step (no synchronisation).  [B200] B = 512: 1.09 ms on the device, 0.41 ms of host
time; B = 4096: 6.68 ms / 0.44 ms -- the host is not the limiter (no CUDA graph needed), the flanger's slowest example is.
    python scripts/strong_probe.py [B]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mod_extraction_b200.render import InterwovenRenderer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cx = bench.Ctx(argparse.Namespace(no_numa_bind=False))
R = InterwovenRenderer(bench.N, float(bench.SR), cx.dev)
b = bench.make_batch4(cx, R, 4096, 43, 0, B)
step = lambda: R.render(b["dry"], b["effect"], b["mod_lo"], b["fc"], b["ph"], wet=b["wet"], logmel=b["logmel"],
                        ph_long=b["ph_long"], ph_start=b["ph_start"])
for _ in range(5):
    step()
torch.cuda.synchronize()
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K):
    step()
t_host = (time.perf_counter() - t0) / K
e1.record(); torch.cuda.synchronize()
print(f"B = {B}: {e0.elapsed_time(e1) / K:.3f} ms per step on the device clock, {t_host * 1e3:.3f} ms of host time to queue one")
