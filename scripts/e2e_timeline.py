"""Device-side timeline of one host-buffer step of bench.py's config 4 (the `e2e` number): how busy the two copy
engines and the SMs are, when each starts and ends, and the largest idle gaps of each engine.  Uses torch.profiler (CUPTI)
on the bench's own e2e step.     python scripts/e2e_timeline.py [B] [chunk]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench
from mod_extraction_b200.modulations import make_combined_mod_sig_batch
from mod_extraction_b200.render import InterwovenRenderer
from mod_extraction_b200.sharding import bind_to_gpu_numa_node

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
import argparse
cx = bench.Ctx(argparse.Namespace(no_numa_bind=False))
dev = cx.dev
N, SR = bench.N, bench.SR
R = InterwovenRenderer(N, float(SR), dev)
wb = bench.make_batch4(cx, R, B, 43, 0, B)
step, info = bench.make_e2e_step(cx, R, wb, chunk)
print("bytes per step: H2D %.3f GB, D2H %.3f GB" % (info["h2d"] / 1e9, info["d2h"] / 1e9))
for _ in range(3):
    step()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
print("wall per step (no profiler): median %.1f ms, min %.1f ms" % (np.median(ts) * 1e3, min(ts) * 1e3))


def burst(n):
    h = None
    for i in range(n):
        nxt = step(wait=False, alt=i & 1)
        if h is not None:
            h.wait()
        h = nxt
    h.wait()


burst(2)
t0 = time.perf_counter(); burst(10); print("pipelined steps (render_host(wait=False)): %.1f ms per step" % ((time.perf_counter() - t0) * 100))

ts = []
for _ in range(5):
    t0 = time.perf_counter(); torch.manual_seed(43)
    m = make_combined_mod_sig_batch(N // 100, SR // 100, wb["rate"], wb["phase"], bench.SHAPES6, device=dev)
    t1 = time.perf_counter(); torch.cuda.synchronize(); ts.append((t1 - t0, time.perf_counter() - t0))
print("LFO synthesis alone: host returns after %.2f ms, device done after %.2f ms" %
      (np.median([a for a, _ in ts]) * 1e3, np.median([b for _, b in ts]) * 1e3))

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter(); step(); wall = time.perf_counter() - t0
print("wall under the profiler: %.1f ms" % (wall * 1e3))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
if not ev:
    sys.exit("no device events recorded")
t_first = min(e.time_range.start for e in ev)


def lane(pred, name):
    xs = sorted((e.time_range.start - t_first, e.time_range.end - t_first) for e in ev if pred(e.name))
    if not xs:
        print(f"{name}: none"); return
    busy = sum(b - a for a, b in xs)
    # merge overlapping intervals for the gap list
    merged = [list(xs[0])]
    for a, b in xs[1:]:
        if a <= merged[-1][1]: merged[-1][1] = max(merged[-1][1], b)
        else: merged.append([a, b])
    gaps = sorted(((merged[i + 1][0] - merged[i][1], merged[i][1]) for i in range(len(merged) - 1)), reverse=True)[:5]
    print(f"{name}: {len(xs)} ops, busy {busy / 1e3:.1f} ms (union {sum(b - a for a, b in merged) / 1e3:.1f}), first start "
          f"{xs[0][0] / 1e3:.2f} ms, last end {max(b for _, b in xs) / 1e3:.2f} ms; largest gaps (ms @ ms): " +
          ", ".join(f"{g / 1e3:.2f}@{at / 1e3:.1f}" for g, at in gaps))


lane(lambda n: "Memcpy HtoD" in n, "H2D")
lane(lambda n: "Memcpy DtoH" in n, "D2H")
lane(lambda n: "Memcpy" not in n and "Memset" not in n, "kernels")
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU]
c0 = min(e.time_range.start for e in cpu)
print("first device op starts %.2f ms after the first host op; host ops span %.1f ms" %
      ((t_first - c0) / 1e3, (max(e.time_range.end for e in cpu) - c0) / 1e3))
big = sorted(((e.time_range.end - e.time_range.start, e.name) for e in ev if "Memcpy" in e.name), reverse=True)[:3]
for d, n in big:
    print("  longest copy: %.2f ms %s" % (d / 1e3, n))

if os.environ.get("E2E_CPROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        step()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
