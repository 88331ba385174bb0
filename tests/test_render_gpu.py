"""GPU: the interwoven batch renderer (config-4 data path) against the oracle, concurrent == serial."""
import numpy as np
import pytest
import torch

import bench

pytestmark = pytest.mark.gpu


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_renderer_concurrent_equals_serial_and_oracle():
    from mod_extraction_b200.render import InterwovenRenderer
    dev = torch.device("cuda", 0)
    B = 24
    dry, effect, mod_lo, fc, ph = bench.oracle_inputs(B, seed=11)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    args = (to(dry), torch.from_numpy(effect), to(mod_lo), {k: to(v) for k, v in fc.items()},
            {k: to(v) for k, v in ph.items()})
    wa, la = InterwovenRenderer(bench.N, float(bench.SR), dev, concurrent=True).render(*args)
    wb, lb = InterwovenRenderer(bench.N, float(bench.SR), dev, concurrent=False).render(*args)
    torch.cuda.synchronize()
    assert torch.equal(wa, wb) and torch.equal(la, lb)
    from mod_extraction_b200.models import LogMelSpectrogram
    ref_wet, ref_lm = bench.oracle_step(dry, effect, mod_lo, fc, ph, threads=4, fb=LogMelSpectrogram().fb.numpy())
    fcx = np.nonzero(effect != 2)[0]
    assert np.array_equal(wa.cpu().numpy()[fcx], ref_wet[fcx])
    # log-mel of rows whose wet audio is bit-identical on both sides (dry halves, flanger / chorus rows): white noise,
    # float32 FFTs on both sides, each <= ~2e-4 from exact arithmetic
    err = np.abs(la.cpu().numpy() - ref_lm)[fcx]
    assert err.max() <= 5e-4 and (err <= 1e-4).mean() >= 0.9999, float(err.max())
    # effect ids outside {0, 1, 2} are rejected instead of leaving rows unwritten
    bad = torch.from_numpy(effect.copy())
    bad[3] = 7
    with pytest.raises(ValueError):
        InterwovenRenderer(bench.N, float(bench.SR), dev).render(args[0], bad, *args[2:])


def test_renderer_index_lists_follow_in_place_refills():
    """One `effect` buffer refilled in place between steps (the pinned-buffer pattern) must not get stale index lists."""
    from mod_extraction_b200.render import InterwovenRenderer
    dev = torch.device("cuda", 0)
    B = 6
    dry, effect, mod_lo, fc, ph = bench.oracle_inputs(B, seed=3)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    R = InterwovenRenderer(bench.N, float(bench.SR), dev)
    eff = torch.from_numpy(effect.copy())
    rest = (to(mod_lo), {k: to(v) for k, v in fc.items()}, {k: to(v) for k, v in ph.items()})
    w1, _ = R.render(to(dry), eff, *rest)
    w1 = w1.clone()
    eff.copy_(torch.tensor([1, 1, 0, 0, 2, 2]))                 # same tensor object, new contents
    w2, _ = R.render(to(dry), eff, *rest)
    fresh, _ = InterwovenRenderer(bench.N, float(bench.SR), dev).render(to(dry), eff.clone(), *rest)
    torch.cuda.synchronize()
    assert torch.equal(w2, fresh) and not torch.equal(w1, w2)


def test_render_host_pipelined_equals_device_render():
    """Host-buffer entry point (pinned in / pinned out, chunked pipeline) == one device-resident render."""
    from mod_extraction_b200.render import InterwovenRenderer
    dev = torch.device("cuda", 0)
    B = 20
    dry, effect, mod_lo, fc, ph = bench.oracle_inputs(B, seed=5)
    R = InterwovenRenderer(bench.N, float(bench.SR), dev)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    eff = torch.from_numpy(effect)
    wet_ref, lm_ref = R.render(to(dry), eff, to(mod_lo), {k: to(v) for k, v in fc.items()},
                               {k: to(v) for k, v in ph.items()})
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    wet_h = torch.empty((B, 1, bench.N)).pin_memory()
    stat_h = torch.empty((B, 2)).pin_memory()
    _, lm = R.alloc_outputs(B)
    R.render_host(pin(dry), eff, pin(mod_lo), {k: pin(v) for k, v in fc.items()}, {k: pin(v) for k, v in ph.items()},
                  wet_h, lm, stat_h, chunk=7)
    assert torch.equal(wet_h, wet_ref.cpu())
    assert torch.equal(lm, lm_ref)
    assert torch.allclose(stat_h, lm_ref.mean(dim=(2, 3)).cpu(), atol=1e-5)


def test_renderer_phaser_rows_from_long_chunks():
    """datasets.py:428-447 inside the interleaved batch: the phaser examples are rendered over N + one LFO period and a
    window of N samples of wet AND dry is kept; the other rows are untouched by it."""
    from oracle import oracle
    from mod_extraction_b200.models import LogMelSpectrogram
    from mod_extraction_b200.render import InterwovenRenderer
    dev = torch.device("cuda", 0)
    B = 12
    dry, effect, mod_lo, fc, ph = bench.oracle_inputs(B, seed=21)
    rng = np.random.RandomState(3)
    phx = np.nonzero(effect == 2)[0]
    extra = (bench.SR / ph["rate_hz"] + 0.5).astype(np.int64)                         # datasets.py:433
    L = bench.N + int(extra[phx].max())
    long_rows = ((rng.random_sample((phx.size, L)) * 2 - 1) * 0.5).astype(np.float32)
    start = np.zeros(B, dtype=np.int32)
    start[phx] = [rng.randint(0, extra[b] + 1) for b in phx]                          # datasets.py:445
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    fcd, phd = {k: to(v) for k, v in fc.items()}, {k: to(v) for k, v in ph.items()}
    outs = []
    for concurrent in (True, False):
        d = to(dry)
        R = InterwovenRenderer(bench.N, float(bench.SR), dev, concurrent=concurrent)
        w, lm = R.render(d, torch.from_numpy(effect), to(mod_lo), fcd, phd, ph_long=to(long_rows), ph_start=to(start))
        torch.cuda.synchronize()
        outs.append((d.cpu().numpy(), w.cpu().numpy(), lm.cpu().numpy()))
    for a, b_ in zip(outs[0], outs[1]):
        assert np.array_equal(a, b_)
    d, w, lm = outs[0]
    fb = LogMelSpectrogram().fb.numpy()
    for k, b in enumerate(phx):
        proc_n = bench.N + int(extra[b])
        ref = oracle.phaser(long_rows[k:k + 1, :proc_n], float(bench.SR), *[ph[n][b:b + 1] for n in bench.PH_KEYS])[0]
        s0 = int(start[b])
        assert np.array_equal(d[b, 0], long_rows[k, s0:s0 + bench.N])                 # dry window, in place
        assert np.abs(w[b, 0] - ref[s0:s0 + bench.N]).max() <= 1e-4
        tru = oracle.log_mel(np.stack([d[b, 0], w[b, 0]]), fb=fb, fft_dtype=np.float64)
        assert (np.abs(lm[b] - tru) <= 1e-4).mean() >= 0.9999
    others = np.nonzero(effect != 2)[0]
    assert np.array_equal(d[others], dry[others])
    plain_w, _ = InterwovenRenderer(bench.N, float(bench.SR), dev).render(to(dry), torch.from_numpy(effect), to(mod_lo), fcd, phd)
    assert np.array_equal(w[others], plain_w.cpu().numpy()[others])


def test_render_host_grouped_buffers_equal_device_render():
    """Host-buffer entry point with the per-effect layout: compact dry rows of the flanger / chorus examples + the longer
    chunks of the phaser examples in, wet audio + the dry windows of the phaser examples out == one device render."""
    from mod_extraction_b200.render import InterwovenRenderer
    dev = torch.device("cuda", 0)
    B = 26
    dry, effect, mod_lo, fc, ph = bench.oracle_inputs(B, seed=9)
    rng = np.random.RandomState(4)
    phx, fcx = np.nonzero(effect == 2)[0], np.nonzero(effect != 2)[0]
    extra = (bench.SR / ph["rate_hz"] + 0.5).astype(np.int64)
    L = bench.N + int(extra[phx].max())
    long_rows = ((rng.random_sample((phx.size, L)) * 2 - 1) * 0.5).astype(np.float32)
    start = np.zeros(B, dtype=np.int32)
    start[phx] = [rng.randint(0, extra[b] + 1) for b in phx]
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    R = InterwovenRenderer(bench.N, float(bench.SR), dev)
    eff = torch.from_numpy(effect)
    d_ref = to(dry)
    w_ref, lm_ref = R.render(d_ref, eff, to(mod_lo), {k: to(v) for k, v in fc.items()}, {k: to(v) for k, v in ph.items()},
                             ph_long=to(long_rows), ph_start=to(start))
    torch.cuda.synchronize()
    wet_h = torch.empty((B, 1, bench.N)).pin_memory()
    dry_ph_h = torch.empty((phx.size, bench.N)).pin_memory()
    stat_h = torch.empty((B, 2)).pin_memory()
    _, lm = R.alloc_outputs(B)
    R.render_host(torch.empty((B, 1, bench.N), device="meta"), eff, pin(mod_lo), {k: pin(v) for k, v in fc.items()},
                  {k: pin(v) for k, v in ph.items()}, wet_h, lm, stat_h, chunk=7, ph_long_h=pin(long_rows),
                  ph_start_h=pin(start), dry_ph_h=dry_ph_h, dry_fc_h=pin(dry[fcx, 0]))
    assert torch.equal(wet_h, w_ref.cpu()) and torch.equal(lm, lm_ref)
    assert torch.equal(dry_ph_h, d_ref.cpu()[phx, 0])
    assert torch.allclose(stat_h, lm_ref.mean(dim=(2, 3)).cpu(), atol=1e-5)


def test_render_host_packed_phaser_rows_and_lfo_callable():
    """Host-buffer entry point fed like a collate function would: the phaser chunks packed back to back, each cut to the
    start + N samples that determine its window (garbage behind them must not matter), the LFOs produced by a callable on
    the device.  Result == one device render of the padded rows, bit for bit; the dry windows only when asked for."""
    from mod_extraction_b200.render import InterwovenRenderer
    dev = torch.device("cuda", 0)
    B = 26
    dry, effect, mod_lo, fc, ph = bench.oracle_inputs(B, seed=11)
    rng = np.random.RandomState(5)
    phx, fcx = np.nonzero(effect == 2)[0], np.nonzero(effect != 2)[0]
    extra = (bench.SR / ph["rate_hz"] + 0.5).astype(np.int64)
    L = bench.N + int(extra[phx].max())
    long_rows = ((rng.random_sample((phx.size, L)) * 2 - 1) * 0.5).astype(np.float32)
    start = np.zeros(B, dtype=np.int32)
    start[phx] = [rng.randint(0, extra[b] + 1) for b in phx]
    start[phx[0]], start[phx[-1]] = 0, extra[phx[-1]]                     # shortest and longest possible prefix
    lens = (start[phx].astype(np.int64) + bench.N + 3) // 4 * 4
    offs = np.concatenate([[0], np.cumsum(lens)])
    packed = np.full((int(offs[-1]),), np.nan, dtype=np.float32)
    for i in range(phx.size):
        n = int(start[phx[i]]) + bench.N
        packed[offs[i]:offs[i] + n] = long_rows[i, :n]                    # the alignment padding stays NaN: never read
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    R = InterwovenRenderer(bench.N, float(bench.SR), dev)
    eff = torch.from_numpy(effect)
    d_ref = to(dry)
    w_ref, lm_ref = R.render(d_ref, eff, to(mod_lo), {k: to(v) for k, v in fc.items()}, {k: to(v) for k, v in ph.items()},
                             ph_long=to(long_rows), ph_start=to(start))
    torch.cuda.synchronize()
    for want_dry, chunk in ((False, 5), (True, 5), (False, 2)):       # chunk 2: some chunks hold no phaser example at all
        wet_h = torch.empty((B, 1, bench.N)).pin_memory()
        dry_ph_h = torch.empty((phx.size, bench.N)).pin_memory() if want_dry else None
        _, lm = R.alloc_outputs(B)
        calls = []

        def lfos():
            calls.append(1)
            return to(mod_lo)
        R.render_host(torch.empty((B, 1, bench.N), device="meta"), eff, lfos, {k: pin(v) for k, v in fc.items()},
                      {k: pin(v) for k, v in ph.items()}, wet_h, lm, None, chunk=chunk, ph_packed_h=pin(packed), ph_offsets=offs,
                      ph_start_h=pin(start), dry_ph_h=dry_ph_h, dry_fc_h=pin(dry[fcx, 0]),
                      duplex=not want_dry)                                # both copy schedules give the same bytes
        assert len(calls) == 1
        assert torch.equal(wet_h, w_ref.cpu()) and torch.equal(lm, lm_ref)
        if want_dry:
            assert torch.equal(dry_ph_h, d_ref.cpu()[phx, 0])


def test_render_host_pipelined_steps_equal_blocking_steps():
    """render_host(wait=False): three consecutive steps with different inputs share the device buffers and overlap; each
    lands exactly what its own blocking call produces (per-chunk events order the reuse of dry_d / wet_d / logmel)."""
    from mod_extraction_b200.render import InterwovenRenderer
    dev = torch.device("cuda", 0)
    B = 23
    R = InterwovenRenderer(bench.N, float(bench.SR), dev)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    steps = []
    for seed in (21, 22, 23):
        dry, effect, mod_lo, fc, ph = bench.oracle_inputs(B, seed=seed)
        steps.append((pin(dry), torch.from_numpy(effect), pin(mod_lo), {k: pin(v) for k, v in fc.items()},
                      {k: pin(v) for k, v in ph.items()}))
    dry_d = torch.empty((B, 1, bench.N), device=dev)
    wet_d, lm = R.alloc_outputs(B)
    ref = []
    for d, e, m, f, p in steps:
        w = torch.empty((B, 1, bench.N)).pin_memory()
        st = torch.empty((B, 2)).pin_memory()
        R.render_host(d, e, m, f, p, w, lm, st, chunk=6, dry_d=dry_d, wet_d=wet_d)
        ref.append((w.clone(), st.clone()))
    outs, handles = [], []
    for k, (d, e, m, f, p) in enumerate(steps):
        w = torch.empty((B, 1, bench.N)).pin_memory()
        st = torch.empty((B, 2)).pin_memory()
        # the last step is cut into other chunks than its predecessor: ordered behind the whole previous step instead
        handles.append(R.render_host(d, e, m, f, p, w, lm, st, chunk=6 if k < 2 else 4, dry_d=dry_d, wet_d=wet_d, wait=False))
        outs.append((w, st))
    for h in handles:
        h.wait()
        h.wait()                                                          # idempotent
    for (w, st), (rw, rst) in zip(outs, ref):
        # (the per-example log-mel mean is a torch reduction whose summation order follows the chunk shape: last bits)
        assert torch.equal(w, rw) and torch.allclose(st, rst, rtol=0, atol=1e-5)
