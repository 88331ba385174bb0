import torch, time
dev = "cuda:0"
n = 1 << 28   # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
gb = n * 4 / 1e9
th = t(lambda: d_in.copy_(h_in, non_blocking=True)); print(f"H2D {gb/th:.1f} GB/s")
td = t(lambda: h_out.copy_(d_out, non_blocking=True)); print(f"D2H {gb/td:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
tb = t(both); print(f"H2D+D2H concurrently: {gb/tb:.1f} GB/s each direction ({2*gb/tb:.1f} total)")
