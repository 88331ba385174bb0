import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mod_extraction_b200.models import LogMelSpectrogram
dev = "cuda:0"
front = LogMelSpectrogram().to(dev)
for B in (64, 683, 2048, 4096):
    x = (torch.rand((B, 2, 88200), device=dev) - 0.5)
    out = torch.empty((B, 2, 256, 345), device=dev)
    for _ in range(3): front(x, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); front(x, out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    byt = B * 2 * 88200 * 4 + out.numel() * 4
    print(f"logmel B={B:5d}: {t:8.3f} ms  {B*2.0/(t*1e-3)/1e6:7.3f} M audio-s/s  {byt/(t*1e-3)/1e9:7.1f} GB/s ({byt/(t*1e-3)/1e9/6545*100:5.1f}% of 6545)")
