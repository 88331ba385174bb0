"""Per-layer timing of the LFO-net body (N3): python scripts/quick_cnn_bench.py [B] [precision ...]
Prints ms and TFLOP/s of every conv + pool + PReLU launch, the layer norms, and the whole extractor."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod_extraction_b200 import _lib                                     # noqa: E402
from mod_extraction_b200.models import Spectral2DCNN, _stream, _vp, round_to_tf32       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
precisions = sys.argv[2:] or ["tf32", "fp32"]
dev = "cuda:0"
L = _lib.lib()
DIL = [1, 1, 2, 4, 8, 16]


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))[n // 2]


W = 345
for prec in precisions:
    code = _lib.CNN_TF32 if prec == "tf32" else _lib.CNN_FP32
    total = 0.0
    for i in range(1, 6):
        H = 256 >> i
        x = round_to_tf32(torch.randn(B, H, W, 64, device=dev))
        w = round_to_tf32(torch.randn(5, 13, 64, 64, device=dev) * 0.02)
        b = torch.zeros(64, device=dev)
        y = torch.empty(B, H // 2, W, 64, device=dev)
        ms = timed(lambda: _lib.check(L.modfx_cnn_conv_pool_prelu_f32(_vp(x), _vp(y), B, H, W, 64, 64, 5, 13, DIL[i], _vp(w),
                                                                      _vp(b), _vp(b), code, _stream())))
        fl = 2.0 * B * H * W * 64 * 64 * 65
        total += ms
        print(f"{prec} layer {i + 1}: H={H:3d} dil={DIL[i]:2d}  {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s")
    print(f"{prec} layers 2-6: {total:.3f} ms for B={B} ({total / B * 1e3:.1f} us per example)")

if "tf32" in precisions:
    total = 0.0
    for i in range(1, 6):
        H = 256 >> i
        xh = torch.randn(B, H, W, 64, device=dev).half()
        wh = (torch.randn(5, 13, 64, 64, device=dev) * 0.02).half()
        b = torch.zeros(64, device=dev)
        y = torch.empty(B, H // 2, W, 64, device=dev)
        ms = timed(lambda: _lib.check(L.modfx_cnn_conv_pool_prelu_f16_f32(_vp(xh), _vp(y), B, H, W, DIL[i], _vp(wh), _vp(b), _vp(b),
                                                                          _stream())))
        total += ms
        print(f"fp16 layer {i + 1}: H={H:3d} dil={DIL[i]:2d}  {ms:8.3f} ms  {2.0 * B * H * W * 64 * 64 * 65 / ms / 1e9:7.1f} TFLOP/s")
    print(f"fp16 layers 2-6: {total:.3f} ms for B={B} ({total / B * 1e3:.1f} us per example)")

x = torch.randn(B, 256, W, 2, device=dev)
w = torch.randn(5, 13, 64, 2, device=dev) * 0.1
b = torch.zeros(64, device=dev)
y = torch.empty(B, 128, W, 64, device=dev)
for code, name in ((0, "CUDA cores"), (1, "tcgen05")):
    ms = timed(lambda: _lib.check(L.modfx_cnn_conv_pool_prelu_f32(_vp(x), _vp(y), B, 256, W, 2, 64, 5, 13, 1, _vp(w), _vp(b), _vp(b),
                                                                  code, _stream())))
    print(f"layer 1 (2 -> 64, {name}): {ms:.3f} ms  {2.0 * B * 256 * W * 2 * 64 * 65 / ms / 1e9:.1f} TFLOP/s  "
          f"{y.numel() * 4 / ms / 1e6:.0f} GB/s of output")
ws = torch.empty(L.modfx_cnn_layernorm_workspace_bytes(B, 64, 128, W), dtype=torch.uint8, device=dev)
ms = timed(lambda: _lib.check(L.modfx_cnn_layernorm_f32(_vp(y), _vp(y), B, 64, 128, W, 0, 1e-5, 1, _vp(ws), _stream())))
print(f"layer norm of (B, 128, 345, 64): {ms:.3f} ms  {2 * y.numel() * 4 / ms / 1e6:.0f} GB/s (read twice + write once: x1.5)")

for prec in precisions + (["fp16", "tf32x3"] if "tf32" in precisions else []):
    net = Spectral2DCNN(in_ch=2, out_channels=[64] * 6, temp_dilations=DIL, pool_size=(2, 1), precision=prec).to(dev).eval()
    audio = torch.rand(B, 2, 88200, device=dev) - 0.5
    ms = timed(lambda: net(audio), n=3)
    print(f"{prec} extractor (log-mel + CNN), B={B}: {ms:.2f} ms  -> {B * 2.0 / (ms / 1e3):.0f} audio-s/s")
