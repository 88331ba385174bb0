import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mod_extraction_b200.phaser import Phaser
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1365
N = 88200
dev = "cuda:0"
x = (torch.rand((B, N), device=dev) - 0.5)
rng = np.random.RandomState(0)
logu = lambda lo, hi: torch.from_numpy(np.exp(rng.uniform(np.log(lo), np.log(hi), B)).astype(np.float32))
U = lambda lo, hi: torch.from_numpy(rng.uniform(lo, hi, B).astype(np.float32))
ps = [logu(0.5, 3.0), U(0.2, 1.0), logu(70.0, 18000.0), U(0.0, 0.7), U(0.2, 1.0)]
ph = Phaser(44100.0)
out = torch.empty_like(x)
for _ in range(3):
    ph(x, *ps, out=out)
torch.cuda.synchronize()
