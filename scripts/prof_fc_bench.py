"""Flanger (or chorus) launch with the bench's config-4 inputs, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mod_extraction_b200 import _ops
from mod_extraction_b200._ops import ModSource
from mod_extraction_b200.modulations import make_combined_mod_sig_batch
from mod_extraction_b200.render import InterwovenRenderer
which = sys.argv[1] if len(sys.argv) > 1 else "flanger"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda", 0)
effect_np, fc_np, ph_np, rate, phase = bench.host_params(B, 43)
torch.manual_seed(43)
dry = (torch.rand((B, 1, bench.N), device=dev) * 2 - 1) * 0.5
mod_lo = make_combined_mod_sig_batch(bench.N // 100, bench.SR // 100, rate, phase, bench.SHAPES6, device=dev)
fc = [torch.from_numpy(fc_np[k]).to(dev) for k in ("feedback", "min_delay_width", "width", "depth", "mix")]
R = InterwovenRenderer(bench.N, float(bench.SR), dev, concurrent=False)
i_fl, i_ch, i_ph, _ = R._groups(torch.from_numpy(effect_np))
wet = torch.empty_like(dry)
idx, dl = (i_fl, R.fl) if which == "flanger" else (i_ch, R.ch)
for _ in range(3):
    _ops.flanger_chorus(dry, ModSource.control_rate(mod_lo), dl[0], dl[1], *fc, example_index=idx, out=wet)
torch.cuda.synchronize()
