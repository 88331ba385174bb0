"""Schedule statistics of fc_cta_kernel's consumer warp (needs the -DMODFX_FC_STATS build: scripts/build_variant.sh):
per example, how many blocks went through each schedule and how many cycles they took.
    MODFX_LIB=gpurun_out/libmodfx_stats.so python scripts/fc_stats.py [B]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mod_extraction_b200.fx import MonoFlangerChorusModule
from mod_extraction_b200.modulations import make_combined_mod_sig_batch, make_mod_signal_batch
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 171
kind = sys.argv[2] if len(sys.argv) > 2 else "combined"
N, SR = 88200, 44100
SH = ["cos", "tri", "rect_cos", "inv_rect_cos", "saw", "rsaw"]
rng = np.random.RandomState(44); torch.manual_seed(44)
if kind == "same":      # B copies of BASELINE config 1 (one 2 s clip, 2 Hz triangle): one CTA per SM when B = 148
    lo = make_mod_signal_batch(882, 441.0, [2.0] * B, [0.0] * B, ["tri"] * B)
elif kind == "combined":
    lo = make_combined_mod_sig_batch(882, 441, np.exp(rng.uniform(0.0, math.log(3.0), B)), rng.uniform(0, 2 * math.pi, B), SH, dev)
else:
    lo = make_mod_signal_batch(882, 441.0, np.exp(rng.uniform(math.log(0.5), math.log(3.0), B)), rng.uniform(0, 2 * math.pi, B),
                               [SH[b % 6] for b in range(B)], np.full(B, 2.0))
g = torch.Generator(device=dev).manual_seed(1)
x = (torch.rand((B, 1, N), device=dev, generator=g) * 2 - 1) * 0.5
U = lambda a, b: torch.from_numpy(rng.uniform(a, b, B).astype(np.float32)).to(dev)
p = [U(0, 0.7), U(0, 1.0), U(0.25, 1), U(0.25, 1), U(0.25, 1)]
if kind == "same":
    p = [torch.full((B,), v, device=dev) for v in (0.5, 0.1, 1.0, 1.0, 1.0)]
m = MonoFlangerChorusModule(B, 1, N, SR, 1.0, 10.0, check_ranges=False)
stats = torch.zeros((B * 16,), dtype=torch.int64, device=dev)
os.environ["MODFX_FC_STATS_PTR"] = str(stats.data_ptr())
os.environ["MODFX_FC_KERNEL"] = "cta"
out = torch.empty_like(x)
for _ in range(2):
    m.forward_control_rate(x, lo, *p, out=out)
torch.cuda.synchronize()
sall = stats.cpu().numpy().astype(np.float64)
s = sall[:B * 12].reshape(B, 12)
pr = sall[B * 12:].reshape(B, 4)
nt = (N + 127) // 128
print('producer 0, mean cycles per tile it handled: wait-for-consumer %.0f  epilogue %.0f  audio-wait %.0f  fill %.0f' % tuple(pr.mean(0) / (nt / 3)))
names = ["wait", "one-shot(tiles)", "one-wave", "serial", "waves", "generic"]
tot = s[:, 6:].sum(1)
order = np.argsort(-tot)
print("per-example consumer cycles: mean %.0f  max %.0f  (%.3f ms at 1.965 GHz)" % (tot.mean(), tot.max(), tot.max() / 1.965e6))
for title, rows in (("mean over examples", None), ("slowest example", order[0]), ("2nd slowest", order[1])):
    r = s.mean(0) if rows is None else s[rows]
    print(f"-- {title}: total {r[6:].sum():.0f} cycles")
    for k, nm in enumerate(names):
        cnt, cyc = r[k], r[6 + k]
        unit = 128 if k == 1 else 32
        print(f"   {nm:16s} count {cnt:9.1f}  cycles {cyc:10.0f} ({cyc / max(r[6:].sum(), 1) * 100:5.1f} %)  "
              f"{cyc / max(cnt, 1):8.1f} cyc each" + ("" if k == 0 else f"  {cyc / max(cnt * unit, 1):6.2f} cyc/sample"))
i = order[0]
print("slowest example params: fb %.3f mdw %.3f width %.3f; lo min %.4f max %.4f" % (float(p[0][i]), float(p[1][i]), float(p[2][i]), float(lo[i].min()), float(lo[i].max())))

import ctypes
from mod_extraction_b200 import _lib
L = ctypes.CDLL(_lib.LIB_PATH)
if hasattr(L, "modfx_fc_serial_stats"):
    buf = (ctypes.c_ulonglong * 4)()
    L.modfx_fc_serial_stats(buf)
    if buf[0]:
        print("serial_run: %d calls, %.1f blocks per call, %.0f cycles per call, %.1f cycles per sample inside the loop" % (buf[0], buf[1] / 8.0 / buf[0], buf[2] / buf[0], buf[3] / (4.0 * buf[1])))

# ---- how well fc_cost_kernel's score predicts the measured consumer cycles, and what list scheduling makes of it
A = (441.0 * p[2].cpu().numpy()).astype(np.float32)                 # Mlfo * width
D0 = (44.0 * p[1].cpu().numpy()).astype(np.float32)                 # min_delay_width * Mmin
d = A[:, None] * lo.cpu().numpy() + D0[:, None]
cost = np.where(d < 1.0, 1.0, np.where(d < 9.0, 16.0, np.where(d < 33.0, 1.0 + 60.0 / np.maximum(d - 1.0, 1e-3), np.where(d < 129.0, 1.5, 1.0))))
score = cost.sum(1)
rk = lambda v: np.argsort(np.argsort(v))
print("rank correlation of the predicted cost with the measured consumer cycles: %.3f" % np.corrcoef(rk(score), rk(tot))[0, 1])
lin = np.polyfit(score, tot, 1)
print("measured cycles ~ %.1f * score + %.0f; residual rms %.0f cycles (mean %.0f)" % (lin[0], lin[1], np.sqrt(np.mean((np.polyval(lin, score) - tot) ** 2)), tot.mean()))


def makespan(order_, slots=148 * 7):
    import heapq
    free = [0.0] * slots
    heapq.heapify(free)
    end = 0.0
    for i in order_:
        t = heapq.heappop(free) + tot[i]
        end = max(end, t)
        heapq.heappush(free, t)
    return end / 1.965e6


if B > 148 * 7:
    print("list scheduling on %d slots with the measured per-example cycles: index order %.3f ms, by predicted cost %.3f ms, "
          "by measured cost %.3f ms; sum / slots %.3f ms" % (148 * 7, makespan(range(B)), makespan(np.argsort(-score)),
                                                               makespan(np.argsort(-tot)), tot.sum() / (148 * 7) / 1.965e6))
