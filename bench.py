#!/usr/bin/env python
"""Benchmark of the render hot path (BASELINE.json metric: audio-seconds rendered per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE config 4 (`train_lfo_interwoven_all` data path): examples interleaved
flanger / chorus / phaser (datasets.py:79-83), combined-shape control-rate LFOs for flanger and chorus
(upsampled x100 inside the effect kernel), then the log-mel front end of cat[dry, wet] ->
(B, 2, 256, 345).  One step = one pass over one batch of synthetic 2 s mono 44.1 kHz clips.
Examples are independent, so ranks render disjoint batches with no collective (weak scaling).

Prints ONE JSON line on rank 0 (see the keys at the bottom).  `--impl reference` times the CPU
restatement of the reference algorithm (oracle/, kind "port": the reference itself is python and
cannot travel to the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
N = 88200
N_LO = N // 100
N_MELS, N_FRAMES = 256, N // 256 + 1
SHAPES6 = ["cos", "tri", "rect_cos", "inv_rect_cos", "saw", "rsaw"]
BYTES_PER_EXAMPLE = N * 4 + N * 4 + 2 * N_MELS * N_FRAMES * 4       # read dry, write wet, write log-mel = 1 412 160
WORKLOAD = "config4: interwoven flanger/chorus/phaser + combined control-rate LFO + log-mel (B,2,256,345)"


def env_int(name, default):
    return int(os.environ.get(name, default))


# --------------------------------------------------------------------------------------------- inputs

def host_params(B, seed):
    """Per-example parameters with the reference's ranges (configs/data/gen_idmt_{fl,ch}.yml:34-51,
    configs/data/interwoven_idmt_all.yml:24-40, configs/eval_lfo_combined.yml:35-49)."""
    rng = np.random.RandomState(seed)
    effect = np.arange(B) % 3                                   # flanger, chorus, phaser (datasets.py:79-83)
    U = lambda lo, hi: rng.uniform(lo, hi, B).astype(np.float32)
    logU = lambda lo, hi: np.exp(rng.uniform(math.log(lo), math.log(hi), B))
    mdw = U(0.0, 1.0)
    mdw_ch = U(0.367, 1.0)
    fc = {"feedback": U(0.0, 0.7), "min_delay_width": np.where(effect == 1, mdw_ch, mdw).astype(np.float32),
          "width": U(0.25, 1.0), "depth": U(0.25, 1.0), "mix": U(0.25, 1.0)}
    ph = {"rate_hz": logU(0.5, 3.0).astype(np.float32), "depth": U(0.2, 1.0),
          "centre_frequency_hz": logU(70.0, 18000.0).astype(np.float32), "feedback": U(0.0, 0.7), "mix": U(0.2, 1.0)}
    lfo_rate = logU(1.0, 3.0)
    lfo_phase = rng.uniform(0.0, 2 * math.pi, B)
    return effect, fc, ph, lfo_rate, lfo_phase


# --------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm

def oracle_step(dry, effect, mod_lo, fc, ph, threads):
    """The same workload through the CPU restatement (oracle/): returns (wet, logmel)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    B = dry.shape[0]
    wet = np.empty_like(dry)
    for k, (mmd, mld) in ((0, (1.0, 10.0)), (1, (30.0, 10.0))):
        idx = np.nonzero(effect == k)[0]
        if idx.size:
            mod = oracle.linear_interpolate_last_dim(mod_lo[idx], N)
            wet[idx] = oracle.flanger_chorus(dry[idx], mod, *[fc[n][idx] for n in
                                                              ("feedback", "min_delay_width", "width", "depth", "mix")],
                                             sr=SR, max_min_delay_ms=mmd, max_lfo_delay_ms=mld)
    idx = np.nonzero(effect == 2)[0]
    if idx.size:
        wet[idx, 0] = oracle.phaser(dry[idx, 0], float(SR), *[ph[n][idx] for n in
                                                              ("rate_hz", "depth", "centre_frequency_hz", "feedback", "mix")])
    both = np.concatenate([dry, wet], axis=1)                    # lightning.py:106
    fb = oracle.mel_filterbank()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        rows = list(ex.map(lambda b: oracle.log_mel(both[b], fb=fb), range(B)))
    return wet, np.stack(rows)


def oracle_inputs(B, seed):
    from oracle import oracle
    effect, fc, ph, rate, phase = host_params(B, seed)
    rng = np.random.RandomState(seed + 1)
    dry = ((rng.random_sample((B, 1, N)) * 2 - 1) * 0.5).astype(np.float32)
    draws = oracle.ReplayDraws(choices=rng.randint(0, 6, 64 * B))
    mod_lo = np.stack([oracle.make_combined_mod_sig(N_LO, SR // 100, rate[b], phase[b], SHAPES6, rng=draws)
                       for b in range(B)])
    return dry, effect, mod_lo, fc, ph


def time_oracle(sample_B, reps, seed=1234):
    from oracle import oracle
    oracle.build()
    threads = oracle.num_threads()
    dry, effect, mod_lo, fc, ph = oracle_inputs(sample_B, seed)
    oracle_step(dry[:4], effect[:4], mod_lo[:4], {k: v[:4] for k, v in fc.items()}, {k: v[:4] for k, v in ph.items()},
                threads)                                        # warm caches / lazy init
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle_step(dry, effect, mod_lo, fc, ph, threads)
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return sample_B * (N / SR) / t, t, threads


def run_reference_arm(args, rank):
    if rank != 0:
        return
    sample_B = args.cpu_sample
    ts = []
    from oracle import oracle
    oracle.build()
    threads = oracle.num_threads()
    dry, effect, mod_lo, fc, ph = oracle_inputs(sample_B, 1234)
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle_step(dry, effect, mod_lo, fc, ph, threads)
        if i >= args.warmup:
            ts.append(time.perf_counter() - t0)
    total = sum(ts)
    value = sample_B * len(ts) * (N / SR) / total
    sample = f"{sample_B} examples x 2 s of the same workload per step, {args.steps} steps"
    line = {
        "impl": "reference", "metric": "audio_seconds_rendered_per_second", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_samples": N, "sr": SR, "examples_per_step": sample_B},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="modfx", choices=["modfx", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=4096,
                    help="examples per GPU per step (BASELINE config 4 names 4096; weak scaling keeps it per GPU)")
    ap.add_argument("--cpu-sample", type=int, default=96, help="examples in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extractor", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU-local CPUs")
    ap.add_argument("--gather-rendered", action="store_true",
                    help="N > 1: also time the optional all-gather of the rendered wet batch (outside the timed region)")
    ap.add_argument("--extractor-batch", type=int, default=128, help="clips per forward of the extractor side measurement")
    ap.add_argument("--e2e-chunk", type=int, default=512, help="examples per pipelined chunk of the host-buffer path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    from mod_extraction_b200.sharding import bind_to_gpu_numa_node
    numa_cpus = None if args.no_numa_bind else bind_to_gpu_numa_node(local_rank)    # before any pinned allocation
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from mod_extraction_b200 import _lib
    from mod_extraction_b200.modulations import make_combined_mod_sig_batch
    from mod_extraction_b200.render import InterwovenRenderer
    _lib.lib()                                                  # fail loudly if the extension is missing

    B = args.batch_per_gpu
    seed = 43 + rank                                            # configs/train_lfo_interwoven_all.yml:1
    effect_np, fc_np, ph_np, rate, phase = host_params(B, seed)
    torch.manual_seed(seed)
    gen = torch.Generator(device=dev).manual_seed(seed)
    dry = (torch.rand((B, 1, N), device=dev, generator=gen) * 2 - 1) * 0.5          # family W
    effect = torch.from_numpy(effect_np)
    t0 = time.perf_counter()
    mod_lo = make_combined_mod_sig_batch(N_LO, SR // 100, rate, phase, SHAPES6, device=dev)
    torch.cuda.synchronize()
    lfo_gen_s = time.perf_counter() - t0
    fc = {k: torch.from_numpy(v).to(dev) for k, v in fc_np.items()}
    ph = {k: torch.from_numpy(v).to(dev) for k, v in ph_np.items()}
    R = InterwovenRenderer(N, float(SR), dev)
    wet, logmel = R.alloc_outputs(B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        R.render(dry, effect, mod_lo, fc, ph, wet=wet, logmel=logmel)

    # ---------------- device-resident throughput (`value`)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B * (N / SR) / (ms_per_step * 1e-3)

    # ---------------- per-kernel durations, serialised (same launches, one stream) for the roofline
    Rs = InterwovenRenderer(N, float(SR), dev, concurrent=False)
    from mod_extraction_b200 import _ops
    from mod_extraction_b200._ops import ModSource
    i_fl, i_ch, i_ph, _ = Rs._groups(effect)
    fc_args = [fc[k] for k in ("feedback", "min_delay_width", "width", "depth", "mix")]
    ph_args = [ph[k] for k in ("rate_hz", "depth", "centre_frequency_hz", "feedback", "mix")]
    nm = N_MELS * N_FRAMES
    dry2, wet2 = dry.view(B, N), wet.view(B, N)
    launches = {
        "flanger": (lambda: _ops.flanger_chorus(dry, ModSource.control_rate(mod_lo), Rs.fl[0], Rs.fl[1], *fc_args,
                                                example_index=i_fl, out=wet), i_fl.numel() * N * 8, 1),
        "chorus": (lambda: _ops.flanger_chorus(dry, ModSource.control_rate(mod_lo), Rs.ch[0], Rs.ch[1], *fc_args,
                                               example_index=i_ch, out=wet), i_ch.numel() * N * 8, 2),
        "phaser": (lambda: _ops.phaser(dry2, float(SR), *ph_args, example_index=i_ph, out=wet2),
                   i_ph.numel() * N * 8, 4),
        "logmel": (lambda: Rs.front.forward_rows(dry2, N, B, logmel.view(-1), N, 2 * nm, None),
                   B * (N * 4 + nm * 4), 1),
    }
    kernels = {}
    n_rep = max(3, min(10, args.steps))
    for name, (fn, nbytes, n_launch) in launches.items():
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n_rep):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        kernels[name] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9, "launches": n_launch}
    step()                                                      # leave wet / log-mel consistent again

    peaks, traffic_tab = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    units = {"logmel": ("bytes_per_row", B), "flanger": ("bytes_per_example", i_fl.numel()),
             "chorus": ("bytes_per_example", i_ch.numel()), "phaser": ("bytes_per_example", i_ph.numel())}
    for name, (key, n_units) in units.items():
        t = traffic_tab.get(name, {}).get(key)
        kernels[name]["dram_traffic_bytes"] = None if t is None else t * n_units
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(kernels, key=lambda k: kernels[k]["ms"] * (2 if k == "logmel" else 1))
    roofline = {"bound": "hbm", "kernel": {"logmel": "logmel_kernel (one launch over B dry rows; the wet half is a second identical launch)",
                                           "flanger": "fc_kernel<control-rate> (flanger group)",
                                           "chorus": "fc_wide_kernel (chorus group)",
                                           "phaser": "phaser_{ctl,map,scan,run}_kernel (4 launches)"}[dom],
                "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["gbs"] / peak,
                "traffic": kernels[dom].get("dram_traffic_bytes"), "peak_source": peak_src,
                "traffic_source": "ncu dram__bytes_read+write per unit of work (profiles/traffic.json), scaled to this launch",
                "how": "algorithmic bytes of the launch / median CUDA-event duration, launches serialised on one stream "
                       "after the timed region (inside the timed region the four streams overlap)",
                "pipeline": {"achieved": world * B * BYTES_PER_EXAMPLE / (ms_per_step * 1e-3) / 1e9 / world,
                             "frac": B * BYTES_PER_EXAMPLE / (ms_per_step * 1e-3) / 1e9 / peak,
                             "bytes_per_example": BYTES_PER_EXAMPLE},
                "kernels": kernels}

    # ---------------- end to end through the public API with host buffers (`e2e`)
    e2e = None
    if not args.no_e2e:
        dry_h = torch.empty((B, 1, N), dtype=torch.float32).pin_memory()
        dry_h.copy_(dry.cpu())
        mod_h = mod_lo.cpu().pin_memory()
        fc_h = {k: v.cpu().pin_memory() for k, v in fc.items()}
        ph_h = {k: v.cpu().pin_memory() for k, v in ph.items()}
        wet_h = torch.empty((B, 1, N), dtype=torch.float32).pin_memory()
        stat_h = torch.empty((B, 2), dtype=torch.float32).pin_memory()
        dry_d = torch.empty_like(dry)
        h2d = dry_h.numel() * 4 + mod_h.numel() * 4 + sum(v.numel() * 4 for v in fc_h.values()) + \
            sum(v.numel() * 4 for v in ph_h.values())
        d2h = wet_h.numel() * 4 + stat_h.numel() * 4

        def e2e_step():
            # pinned host buffers in, pinned host wet audio out, chunked so H2D / kernels / D2H overlap
            R.render_host(dry_h, effect, mod_h, fc_h, ph_h, wet_h, logmel, stat_h, chunk=args.e2e_chunk,
                          dry_d=dry_d, wet_d=wet)

        for _ in range(2):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * (N / SR) / dt, "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "chunk": args.e2e_chunk,
               "note": "InterwovenRenderer.render_host: pinned host dry audio + parameters in, wet audio + per-example "
                       "log-mel mean out, chunks pipelined over copy/compute/copy streams; the (B,2,256,345) log-mel "
                       "tensor stays in HBM where the extractor consumes it"}

    checksum = float(wet.double().abs().mean().item())          # of the rendered batch, before any side measurement
    # ---------------- the consumer of the step's log-mel tensor, reported beside the headline, not in it
    extractor = None
    if not args.no_extractor:
        try:
            from mod_extraction_b200.models import Spectral2DCNN
            Bx = min(args.extractor_batch, B)
            feats = logmel[:Bx]
            flops = 2.0 * 65 * 64 * 345 * (256 * 2 + 64 * (128 + 64 + 32 + 16 + 8)) * Bx      # the six convolutions
            bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
            extractor = {"what": "Spectral2DCNN body on the log-mel of this step (6 x {layer norm, 5x13 conv + pool + PReLU on "
                                 "tcgen05}, head), random weights", "batch": Bx, "gpu_launches_per_forward": 19,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops for float16 operands, half of it for TF32 (the driver "
                                        "measures no TF32 figure)"}
            for prec, peak in (("tf32", bf16_peak / 2.0), ("fp16", bf16_peak)):
                torch.manual_seed(1234)                 # the same random weights for both operand formats
                net = Spectral2DCNN(in_ch=2, n_samples=N, sr=SR, out_channels=[64] * 6, temp_dilations=[1, 1, 2, 4, 8, 16],
                                    pool_size=(2, 1), precision=prec).to(dev).eval()
                for _ in range(2):
                    net.forward_features(feats)
                x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                x0.record()
                for _ in range(reps):
                    out_x, _ = net.forward_features(feats)
                x1.record()
                torch.cuda.synchronize()
                ms_x = x0.elapsed_time(x1) / reps
                extractor[prec] = {"ms": ms_x, "audio_s_per_s": Bx * (N / SR) / (ms_x * 1e-3),
                                   "conv_tflops": flops / (ms_x * 1e-3) / 1e12, "peak_tflops": peak,
                                   "frac_of_peak": flops / (ms_x * 1e-3) / 1e12 / peak, "output_mean": float(out_x.mean().item())}
                del net

        except Exception as exc:      # a side measurement must never take the headline line down with it
            extractor = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---------------- final gather of per-rank metrics (the only collective, outside the timed region)
    gather_ms = None
    if world > 1:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cs = torch.tensor([checksum], dtype=torch.float64, device=dev)
        out = torch.empty((world,), dtype=torch.float64, device=dev)
        g0.record()
        dist.all_gather_into_tensor(out, cs)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        checksum = float(out.mean().item())
    rendered_gather = None
    if world > 1 and args.gather_rendered:
        from mod_extraction_b200.sharding import all_gather_rendered
        if extractor is not None:
            torch.cuda.empty_cache()
        all_gather_rendered(wet[:8], 8 * world)                      # NCCL warm-up
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        full = all_gather_rendered(wet, world * B)
        g1.record()
        torch.cuda.synchronize()
        t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = full.numel() * 4
        rendered_gather = {"ms": float(t.item()), "bytes_gathered_per_rank": nbytes,
                           "inbound_gbs_per_rank": nbytes * (world - 1) / world / (float(t.item()) * 1e-3) / 1e9,
                           "note": "sharding.all_gather_rendered(wet): every rank ends with the whole (world*B, 1, N) batch; "
                                   "optional, outside the timed region (SURVEY H7)"}
        del full

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, t, threads = time_oracle(args.cpu_sample, reps=2)
        cpu_baseline = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port", "host_cpus": os.cpu_count(),
                        "sample": f"{args.cpu_sample} examples x 2 s of the same workload, median of 2 runs ({t:.2f} s each)"}

    if rank == 0:
        line = {
            "metric": "audio_seconds_rendered_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "n_samples": N, "sr": SR,
                       "audio": "white noise U(-0.5,0.5)", "lfo": "combined shapes, 882-pt control rate, generated with "
                       "host RNG before the timed region; x100 upsample fused in the effect kernel",
                       "l2": "inputs (1.4 GB dry + 1.4 GB wet + 2.9 GB log-mel per step) far exceed the 126 MB L2",
                       "parallelism": f"batch-sharded x{world}, no collective while rendering"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * 11,      # per step: flanger 1 + chorus 2 (CTA-per-example kernel, then
            # the one-warp kernel that only finds nothing left to do) + phaser 4 + log-mel 4
            "roofline": roofline,
            "cpu_baseline": cpu_baseline, "extractor": extractor, "lfo_generation_s": lfo_gen_s, "metrics_gather_ms": gather_ms,
            "rendered_gather": rendered_gather,
            "host_binding": None if numa_cpus is None else f"rank 0 pinned to {len(numa_cpus)} GPU-local CPUs (NVML affinity)",
            "wet_abs_mean": checksum,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
