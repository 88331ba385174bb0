"""One flanger/chorus launch per configuration, for ncu captures."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mod_extraction_b200.fx import MonoFlangerChorusModule
from mod_extraction_b200.modulations import make_mod_signal_batch
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "chorus"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = 88200
mmd, mld, mdw_lo = (30.0, 10.0, 0.367) if which == "chorus" else (1.0, 10.0, 0.0)
g = torch.Generator(device="cpu").manual_seed(1)
x = ((torch.rand((B, 1, N), generator=g) * 2 - 1) * 0.5).to(dev)
rng = np.random.RandomState(0)
f = np.exp(rng.uniform(np.log(0.5), np.log(3.0), B)); ph = rng.uniform(0, 2*math.pi, B)
shapes = [["cos","tri","rect_cos","inv_rect_cos","saw","rsaw"][b % 6] for b in range(B)]
lo = make_mod_signal_batch(N // 100, 441.0, f, ph, shapes, np.full(B, 2.0))
U = lambda lo_, hi: torch.from_numpy(rng.uniform(lo_, hi, B).astype(np.float32)).to(dev)
p = [U(0, 0.7), U(mdw_lo, 1.0), U(0.25, 1), U(0.25, 1), U(0.25, 1)]
m = MonoFlangerChorusModule(B, 1, N, 44100, mmd, mld, check_ranges=False)
out = torch.empty_like(x)
for _ in range(3):
    m.forward_control_rate(x, lo, *p, out=out)
torch.cuda.synchronize()
