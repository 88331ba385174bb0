"""Times one conv layer: python scripts/time_conv.py [B] [layer] (MODFX_CNN_DEBUG selects diagnostic modes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod_extraction_b200 import _lib                                     # noqa: E402
from mod_extraction_b200.models import _stream, _vp, round_to_tf32       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
layers = [int(a) for a in sys.argv[2:]] or [2]
DIL = [1, 1, 2, 4, 8, 16]
L = _lib.lib()
dev = "cuda:0"
for layer in layers:
    H, W = 256 >> (layer - 1), 345
    x = round_to_tf32(torch.randn(B, H, W, 64, device=dev))
    w = round_to_tf32(torch.randn(5, 13, 64, 64, device=dev) * 0.02)
    b = torch.zeros(64, device=dev)
    y = torch.empty(B, H // 2, W, 64, device=dev)
    f = lambda: _lib.check(L.modfx_cnn_conv_pool_prelu_f32(_vp(x), _vp(y), B, H, W, 64, 64, 5, 13, DIL[layer - 1], _vp(w), _vp(b),
                                                           _vp(b), _lib.CNN_TF32, _stream()))
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    n = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        f()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))[n // 2]
    print(f"dbg={os.environ.get('MODFX_CNN_DEBUG', '0')} layer {layer} B={B}: {ms:.3f} ms  {2.0 * B * H * W * 64 * 64 * 65 / ms / 1e9:.1f} TFLOP/s")
