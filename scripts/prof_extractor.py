"""ncu target: the whole extractor (log-mel + CNN) on B clips.  python scripts/prof_extractor.py [B] [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod_extraction_b200.models import Spectral2DCNN      # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
net = Spectral2DCNN(in_ch=2, out_channels=[64] * 6, temp_dilations=[1, 1, 2, 4, 8, 16], pool_size=(2, 1),
                    precision=prec).to("cuda:0").eval()
audio = torch.rand(B, 2, 88200, device="cuda:0") - 0.5
for _ in range(3):
    y, lat = net(audio)
torch.cuda.synchronize()
print(float(y.mean()))
