"""ncu target: a few launches of one conv + pool + PReLU layer.  python scripts/prof_cnn.py [B] [layer 2..6] [tf32|fp16|fp32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod_extraction_b200 import _lib                                     # noqa: E402
from mod_extraction_b200.models import _stream, _vp, round_to_tf32       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
DIL = [1, 1, 2, 4, 8, 16]
H, W, dev = 256 >> (layer - 1), 345, "cuda:0"
L = _lib.lib()
x = round_to_tf32(torch.randn(B, H, W, 64, device=dev))
w = round_to_tf32(torch.randn(5, 13, 64, 64, device=dev) * 0.02)
b = torch.zeros(64, device=dev)
y = torch.empty(B, H // 2, W, 64, device=dev)
if prec == "fp16":
    xh, wh = x.half(), w.half()
    for _ in range(3):
        _lib.check(L.modfx_cnn_conv_pool_prelu_f16_f32(_vp(xh), _vp(y), B, H, W, DIL[layer - 1], _vp(wh), _vp(b), _vp(b), _stream()))
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(3):
    _lib.check(L.modfx_cnn_conv_pool_prelu_f32(_vp(x), _vp(y), B, H, W, 64, 64, 5, 13, DIL[layer - 1], _vp(w), _vp(b), _vp(b),
                                               _lib.CNN_TF32 if prec == "tf32" else _lib.CNN_FP32, _stream()))
torch.cuda.synchronize()
