"""Where the host-buffer step (InterwovenRenderer.render_host) spends its time: raw pinned-copy rates of this box in each
direction and both at once, and the step with its three stages isolated (copies only / kernels only / full pipeline)
for several chunk sizes.     python scripts/e2e_probe.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mod_extraction_b200.modulations import make_combined_mod_sig_batch
from mod_extraction_b200.render import InterwovenRenderer
from mod_extraction_b200.sharding import bind_to_gpu_numa_node

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rank = int(os.environ.get("LOCAL_RANK", 0))
bind_to_gpu_numa_node(rank)
dev = torch.device("cuda", rank); torch.cuda.set_device(dev)
N, SR = bench.N, bench.SR


def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


# ---- raw copy rates
nb = 1 << 30
h_in, h_out = torch.empty(nb // 4).pin_memory(), torch.empty(nb // 4).pin_memory()
d_a, d_b = torch.empty(nb // 4, device=dev), torch.empty(nb // 4, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
t = wall(lambda: d_a.copy_(h_in, non_blocking=True)); print(f"H2D 1 GiB pinned: {nb / t / 1e9:.1f} GB/s")
t = wall(lambda: h_out.copy_(d_b, non_blocking=True)); print(f"D2H 1 GiB pinned: {nb / t / 1e9:.1f} GB/s")


def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
t = wall(both); print(f"H2D + D2H at once: {nb / t / 1e9:.1f} GB/s per direction")
for mb in (16, 64, 256):
    n = mb * (1 << 20) // 4
    k = nb // 4 // n
    t = wall(lambda: [d_a[i * n:(i + 1) * n].copy_(h_in[i * n:(i + 1) * n], non_blocking=True) for i in range(k)])
    print(f"H2D in {mb} MiB pieces: {nb / t / 1e9:.1f} GB/s")
del h_in, h_out, d_a, d_b

# ---- the step
effect_np, fc_np, ph_np, rate, phase = bench.host_params(B, 43)
eff = torch.from_numpy(effect_np)
R = InterwovenRenderer(N, float(SR), dev)
g = torch.Generator(device=dev).manual_seed(1)
dry = (torch.rand((B, 1, N), device=dev, generator=g) * 2 - 1) * 0.5
i_fc = R._groups(eff)[3]
n_ph = int((effect_np == 2).sum())
L = N + 88200
pin = lambda t: t.cpu().pin_memory()
dry_fc_h = pin(dry.view(B, N).index_select(0, i_fc))
ph_long_h = pin((torch.rand((n_ph, L), device=dev, generator=g) * 2 - 1) * 0.5)
ph_start_h = torch.from_numpy(ph_np["start_idx"]).pin_memory()
fc_h = {k: torch.from_numpy(v).pin_memory() for k, v in fc_np.items()}
ph_h = {k: torch.from_numpy(v).pin_memory() for k, v in ph_np.items() if k != "start_idx"}
torch.manual_seed(43)
mod = make_combined_mod_sig_batch(N // 100, SR // 100, rate, phase, bench.SHAPES6, device=dev)
wet_h = torch.empty((B, 1, N)).pin_memory(); dry_ph_h = torch.empty((n_ph, N)).pin_memory(); stat_h = torch.empty((B, 2)).pin_memory()
wet, lm = R.alloc_outputs(B); dry_d = torch.empty_like(dry)
meta = torch.empty((B, 1, N), device="meta")
h2d = (dry_fc_h.numel() + ph_long_h.numel()) * 4; d2h = (wet_h.numel() + dry_ph_h.numel()) * 4
print(f"step: B = {B}, H2D {h2d / 1e9:.2f} GB, D2H {d2h / 1e9:.2f} GB")
dev_step = lambda: R.render(dry, eff, mod, {k: v.to(dev) for k, v in fc_h.items()}, {k: v.to(dev) for k, v in ph_h.items()},
                            wet=wet, logmel=lm, ph_long=ph_long_d, ph_start=ph_start_d)
ph_long_d, ph_start_d = ph_long_h.to(dev), ph_start_h.to(dev)
t = wall(dev_step); print(f"kernels only (device resident): {t * 1e3:.1f} ms")
t = wall(lambda: (dry_d.view(B, N)[:dry_fc_h.size(0)].copy_(dry_fc_h, non_blocking=True), ph_long_d.copy_(ph_long_h, non_blocking=True)))
print(f"H2D of the step's inputs alone: {t * 1e3:.1f} ms = {h2d / t / 1e9:.1f} GB/s")
t = wall(lambda: (wet_h.copy_(wet, non_blocking=True), dry_ph_h.copy_(dry_d.view(B, N)[:n_ph], non_blocking=True)))
print(f"D2H of the step's outputs alone: {t * 1e3:.1f} ms = {d2h / t / 1e9:.1f} GB/s")
for chunk in (128, 256, 512, 1024):
    t = wall(lambda: R.render_host(meta, eff, mod, fc_h, ph_h, wet_h, lm, stat_h, chunk=chunk, dry_d=dry_d, wet_d=wet,
                                   ph_long_h=ph_long_h, ph_start_h=ph_start_h, dry_ph_h=dry_ph_h, dry_fc_h=dry_fc_h))
    print(f"render_host chunk {chunk}: {t * 1e3:.1f} ms = {B * 2 / t / 1e3:.0f} k audio-s/s")
