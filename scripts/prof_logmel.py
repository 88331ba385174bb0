import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mod_extraction_b200.models import LogMelSpectrogram
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
front = LogMelSpectrogram().to("cuda:0")
x = (torch.rand((B, 2, 88200), device="cuda:0") - 0.5)
out = torch.empty((B, 2, 256, 345), device="cuda:0")
for _ in range(3): front(x, out=out)
torch.cuda.synchronize()
