"""Host-side ceiling of the host-buffer step when N ranks share one host.
Launched like the bench:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port P scripts/e2e_multi_probe.py
Every rank pins 1 GiB in + 1 GiB out, then ALL ranks at once (barrier on both sides): H2D alone, D2H alone, both
directions; per-rank rates and the aggregate are printed by rank 0, with the PCIe counters nvidia-smi dmon sees during
the duplex phase.  Then the same for the bench's own e2e step (one step alone, and steps pipelined)."""
import argparse, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import bench
from mod_extraction_b200.render import InterwovenRenderer

cx = bench.Ctx(argparse.Namespace(no_numa_bind=False))
rank, world, dev = cx.rank, cx.world, cx.dev
nb = 1 << 30
h_in, h_out = torch.empty(nb // 4).pin_memory(), torch.empty(nb // 4).pin_memory()
d_a, d_b = torch.empty(nb // 4, device=dev), torch.empty(nb // 4, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def together(fn, reps=4):
    """Median wall time of fn() + synchronize with every rank starting at the same barrier; max over ranks."""
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        cx.barrier()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    mine = float(np.median(ts))
    if world == 1:
        return [mine]
    out = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(out, torch.tensor([mine], dtype=torch.float64, device=dev))
    return [float(o.item()) for o in out]


def say(name, ts, gb_per_rank):
    if rank == 0:
        rates = [gb_per_rank / t for t in ts]
        print(f"{name}: per rank {min(rates):.1f} .. {max(rates):.1f} GB/s, aggregate {sum(rates):.1f} GB/s "
              f"({world} ranks at once)", flush=True)


def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)


if rank == 0:
    print(f"world {world}; rank 0 bound to {cx.numa_cpus if cx.numa_cpus is None else len(cx.numa_cpus)} CPUs; "
          f"host has {os.cpu_count()} CPUs", flush=True)
say("H2D 1 GiB pinned", together(lambda: d_a.copy_(h_in, non_blocking=True)), nb / 1e9)
say("D2H 1 GiB pinned", together(lambda: h_out.copy_(d_b, non_blocking=True)), nb / 1e9)
dmon = None
if rank == 0:
    try:
        dmon = subprocess.Popen(["nvidia-smi", "dmon", "-s", "t", "-c", "4"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        dmon = None
say("H2D + D2H at once, per direction", together(lambda: [both() for _ in range(4)], reps=6), 4 * nb / 1e9)
if dmon is not None:
    try:
        print("nvidia-smi dmon -s t during the duplex phase (rxpci / txpci in MB/s per GPU):\n" + dmon.communicate(timeout=20)[0], flush=True)
    except Exception:
        dmon.kill()
del h_in, h_out, d_a, d_b

# ---- the bench's own step
N, SR = bench.N, bench.SR
B = int(os.environ.get("PROBE_B", 4096))
R = InterwovenRenderer(N, float(SR), dev)
wb = bench.make_batch4(cx, R, B, 43 + rank, 0, B)
step, info = bench.make_e2e_step(cx, R, wb, 512)
for _ in range(2):
    step()
ts = together(step, reps=4)
if rank == 0:
    gb = (info["h2d"] + info["d2h"]) / 1e9
    print(f"e2e step alone: {max(ts) * 1e3:.1f} ms (slowest rank; fastest {min(ts) * 1e3:.1f}) = "
          f"{world * B * 2 / max(ts) / 1e3:.0f} k audio-s/s; {gb:.2f} GB per rank per step -> "
          f"{world * gb / max(ts):.1f} GB/s of host traffic in total", flush=True)


def burst(n=6):
    h = None
    for i in range(n):
        nxt = step(wait=False, alt=i & 1)
        if h is not None:
            h.wait()
        h = nxt
    h.wait()


ts = together(burst, reps=3)
if rank == 0:
    per = max(ts) / 6
    print(f"e2e steps pipelined: {per * 1e3:.1f} ms per step = {world * B * 2 / per / 1e3:.0f} k audio-s/s; "
          f"{world * (info['h2d'] + info['d2h']) / 1e9 / per:.1f} GB/s of host traffic in total", flush=True)
if world > 1:
    dist.destroy_process_group()
