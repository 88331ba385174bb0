"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports christhetree/mod_extraction from /root/reference (read-only), runs its torch
CPU path on seeded inputs and stores inputs + outputs as small .npz fixtures beside this
script.  /root/reference does not exist on the GPU box, so the tests only ever read the
committed fixtures.  While generating, every case is also run through oracle/oracle.py and
the difference printed, which is how the oracle was pinned.
"""
import math
import os
import sys

import numpy as np
import torch as tr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from mod_extraction import fx as rfx            # noqa: E402
from mod_extraction import modulations as rmod  # noqa: E402
from mod_extraction import util as rutil        # noqa: E402
from mod_extraction import models as rmodels    # noqa: E402

from oracle import oracle                        # noqa: E402

SR = 44100
SHAPES6 = ["cos", "tri", "rect_cos", "inv_rect_cos", "saw", "rsaw"]


class LoggingTorchRng:
    """Draws from the torch global CPU generator in the reference's own way
    (util.py:32-49) and logs the raw draws so tests can replay them without torch RNG."""

    def __init__(self):
        self.u = []
        self.c = []

    def uniform(self, low, high):
        raw = tr.rand(1)
        self.u.append(raw.item())
        return ((raw * (high - low)) + low).item()

    def choice(self, n):
        v = tr.randint(low=0, high=n, size=(1,)).item()
        self.c.append(v)
        return v


def white(shape, seed):
    g = tr.Generator().manual_seed(seed)
    return (tr.rand(shape, generator=g) * 2.0 - 1.0) * 0.5


def guitar(B, N, seed):
    g = tr.Generator().manual_seed(seed)
    t = tr.arange(N, dtype=tr.float64) / SR
    out = []
    for b in range(B):
        f0 = 110.0 * (1.0 + 0.25 * b)
        sig = sum((0.3 / k) * tr.sin(2 * math.pi * f0 * k * t) * tr.exp(-k * t) for k in range(1, 9))
        out.append(sig.float() + 1e-3 * tr.randn(N, generator=g))
    return tr.stack(out, 0).unsqueeze(1)


def report(name, ref, got):
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    d = np.abs(ref - got).max() if ref.size else 0.0
    print(f"  [{name}] oracle vs reference: max|diff| = {d:.3e}  bitwise={np.array_equal(ref, got)}")
    return d


def gen_lfo():
    print("L1 make_mod_signal")
    cases = []
    outs = []
    rng = np.random.RandomState(7)
    cfgs = [(882, 441.0, 2.0, 0.0, "tri", 1.0)]          # SURVEY golden
    for shape in oracle.SHAPES:
        for _ in range(3):
            f = float(np.exp(rng.uniform(np.log(0.5), np.log(3.0))))
            ph = float(rng.uniform(0, 2 * math.pi))
            e = [1.0, 2.0, 0.6][rng.randint(3)]
            cfgs.append((882, 441.0, f, ph, shape, e))
    cfgs += [(4410, 44100.0, 2.7, math.pi / 2, "cos", 1.0), (3000, 172.5, 1.3, -1.0, "rect_cos", 1.0),
             (400, 400.0, 1.0, 0.0, "saw", 1.0), (77, 77.0, 1.0, 0.0, "inv_rect_cos", 1.0)]
    worst = 0.0
    for (n, sr, f, ph, shape, e) in cfgs:
        ref = rmod.make_mod_signal(n, sr, f, ph, shape, e).numpy()
        got = oracle.make_mod_signal(n, sr, f, ph, shape, e)
        worst = max(worst, np.abs(ref - got).max())
        cases.append((n, sr, f, ph, oracle.SHAPE_ID[shape], e))
        outs.append(ref)
    print(f"  oracle vs reference over {len(cfgs)} cases: max|diff| = {worst:.3e}")
    np.savez_compressed(os.path.join(HERE, "lfo.npz"), cases=np.array(cases, dtype=np.float64),
                        **{f"out{i}": o for i, o in enumerate(outs)})


def gen_interp():
    print("U1 linear_interpolate_last_dim")
    g = tr.Generator().manual_seed(11)
    d = {}
    worst = 0.0
    for i, (rows, I, O, ac) in enumerate([(2, 882, 88200, True), (3, 882, 345, True), (2, 17, 53, True),
                                          (1, 300, 123, True), (2, 64, 1000, False), (1, 200, 77, False),
                                          (1, 26460, 264600, True)]):
        x = tr.rand((rows, I), generator=g)
        ref = rutil.linear_interpolate_last_dim(x, O, align_corners=ac).numpy()
        got = oracle.linear_interpolate_last_dim(x.numpy(), O, ac)
        worst = max(worst, report(f"{rows}x{I}->{O} ac={ac}", ref, got))
        d[f"x{i}"] = x.numpy()
        d[f"y{i}"] = ref
        d[f"cfg{i}"] = np.array([rows, I, O, int(ac)])
    d["n"] = np.array(i + 1)
    np.savez_compressed(os.path.join(HERE, "interp.npz"), **d)


def run_ref_fc(x, mod, params, mmd, mld):
    B, C, N = x.shape
    m = rfx.MonoFlangerChorusModule(B, C, N, SR, mmd, mld)
    return m(x, mod, *params).numpy()


def gen_fc():
    print("E1 MonoFlangerChorusModule")
    d = {}
    k = 0

    def add(name, x, mod, params, mmd, mld):
        nonlocal k
        ref = run_ref_fc(x, mod, params, mmd, mld)
        np_params = [p.numpy() if isinstance(p, tr.Tensor) else p for p in params]
        got = oracle.flanger_chorus(x.numpy(), mod.numpy(), *np_params, sr=SR, max_min_delay_ms=mmd,
                                    max_lfo_delay_ms=mld)
        report(name, ref, got)
        d[f"x{k}"] = x.numpy()
        d[f"mod{k}"] = mod.numpy()
        d[f"y{k}"] = ref
        d[f"delays{k}"] = np.array([mmd, mld], dtype=np.float64)
        for j, p in enumerate(params):
            if isinstance(p, tr.Tensor):
                d[f"p{k}_{j}"] = p.numpy()
            else:
                d[f"p{k}_{j}"] = np.array(p, dtype=np.float64)      # 0-d => python float param
        d[f"name{k}"] = np.array(name)
        k += 1

    # BASELINE config 1: flanger, 2 Hz tri LFO at 882 pts upsampled x100, B=1, 2 s clip.
    N = 88200
    lfo = rmod.make_mod_signal(N // 100, SR // 100, 2.0, 0.0, "tri").unsqueeze(0)
    mod = rutil.linear_interpolate_last_dim(lfo, N)
    x = guitar(1, N, 3)
    add("config1_flanger_tri2hz", x, mod, [0.5, 1.0, 1.0, 1.0, 1.0], 1.0, 10.0)

    g = tr.Generator().manual_seed(5)
    B, N = 3, 6000

    def U(lo, hi):
        return tr.rand(B, generator=g) * (hi - lo) + lo

    for name, mmd, mld, mdw_lo in [("flanger_white_tensor", 1.0, 10.0, 0.0), ("chorus_white_tensor", 30.0, 10.0, 0.367),
                                   ("flanger_eval_white_tensor", 1.0, 4.0, 0.0)]:
        x = white((B, 1, N), 21 + k)
        lo = tr.stack([rmod.make_mod_signal(N // 100, SR // 100, f, p, s)
                       for f, p, s in [(2.9, 0.3, "cos"), (1.1, 4.0, "tri"), (2.0, 1.0, "rsaw")]])
        mod = rutil.linear_interpolate_last_dim(lo, N)
        add(name, x, mod, [U(0.0, 0.7), U(mdw_lo, 1.0), U(0.25, 1.0), U(0.25, 1.0), U(0.25, 1.0)], mmd, mld)

    # sub-sample delays / stale tap: min_delay_width 0 and mod touching 0, strong feedback
    x = white((2, 1, 4000), 99)
    mod = tr.stack([rmod.make_mod_signal(4000, SR, 40.0, 0.0, "cos"),
                    tr.clip(rmod.make_mod_signal(4000, SR, 25.0, 1.0, "tri") * 0.01, 0, 1)])
    add("flanger_subsample_delay", x, mod, [tr.tensor([0.69, 0.5]), tr.tensor([0.0, 0.0]),
                                            tr.tensor([1.0, 0.5]), tr.tensor([1.0, 0.9]), tr.tensor([1.0, 0.7])],
        1.0, 10.0)
    # python-float parameters (scalar promotion path), 2 channels, (B,1,N)-shaped mod_sig
    x = white((2, 2, 3000), 123)
    mod = tr.rand((2, 1, 3000), generator=g).expand(-1, 2, -1).contiguous()
    add("chorus_float_params_2ch", x, mod, [0.3, 0.7, 0.3, 0.9, 0.45], 30.0, 10.0)
    x = white((2, 1, 3000), 124)
    mod = tr.rand((2, 3000), generator=g)
    add("flanger_mixed_params_noise_mod", x, mod, [0.6, tr.tensor([0.1, 0.9]), 0.8, tr.tensor([1.0, 0.2]), 1.0],
        1.0, 10.0)
    # all defaults
    add("flanger_defaults", x, mod, [], 1.0, 10.0)
    d["n"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "flanger_chorus.npz"), **d)


def gen_tremolo():
    print("E2 apply_tremolo")
    g = tr.Generator().manual_seed(8)
    x = white((3, 2, 2000), 31)
    mod = tr.rand((3, 2000), generator=g)
    d = {"x": x.numpy(), "mod": mod.numpy()}
    for i, mix in enumerate([1.0, 0.3]):
        ref = rfx.apply_tremolo(x, mod, mix).numpy()
        report(f"tremolo mix={mix}", ref, oracle.tremolo(x.numpy(), mod.numpy(), mix))
        d[f"y{i}"] = ref
        d[f"mix{i}"] = np.array(mix)
    np.savez_compressed(os.path.join(HERE, "tremolo.npz"), **d)


def gen_rng_lfos():
    print("L3/L4 quasi-periodic and combined LFOs (logged host RNG draws)")
    d = {}
    k = 0
    rs = np.random.RandomState(3)
    for i in range(12):
        f = float(np.exp(rs.uniform(np.log(0.5), np.log(2.0))))
        ph = float(rs.uniform(0, 2 * math.pi))
        shape = SHAPES6[i % 6]
        base = rmod.make_mod_signal(882, 441, f, ph, shape)
        tr.manual_seed(100 + i)
        ref = rmod.make_quasi_periodic(base.clone(), 0.10, 0.3333, 0.10, 0.3333, 0.5).numpy()
        tr.manual_seed(100 + i)
        rng = LoggingTorchRng()
        got = oracle.make_quasi_periodic(base.numpy(), 0.10, 0.3333, 0.10, 0.3333, 0.5, rng=rng)
        report(f"quasi {shape} f={f:.3f}", ref, got)
        d[f"q_base{k}"] = base.numpy()
        d[f"q_out{k}"] = ref
        d[f"q_draws{k}"] = np.array(rng.u, dtype=np.float32)
        k += 1
    d["q_n"] = np.array(k)
    d["q_args"] = np.array([0.10, 0.3333, 0.10, 0.3333, 0.5])
    k = 0
    for i in range(12):
        f = float(np.exp(rs.uniform(np.log(1.0), np.log(3.0))))
        ph = float(rs.uniform(0, 2 * math.pi))
        tr.manual_seed(200 + i)
        try:
            ref = rmod.make_combined_mod_sig(882, 441, f, ph, SHAPES6).numpy()
        except AssertionError:
            print("  (reference asserted on a short section; case skipped)")
            continue
        tr.manual_seed(200 + i)
        rng = LoggingTorchRng()
        got = oracle.make_combined_mod_sig(882, 441, f, ph, SHAPES6, rng=rng)
        report(f"combined f={f:.3f}", ref, got)
        d[f"c_args{k}"] = np.array([882, 441, f, ph])
        d[f"c_out{k}"] = ref
        d[f"c_draws{k}"] = np.array(rng.c, dtype=np.int64)
        k += 1
    d["c_n"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "rng_lfos.npz"), **d)


def gen_logmel():
    print("M1 log-mel front end")
    net = rmodels.Spectral2DCNN(in_ch=2, n_samples=88200, sr=SR)
    net.eval()

    def ref_logmel(x):
        with tr.no_grad():
            s = net.spectrogram(x)
            return tr.log(tr.clip(s, min=net.eps)).numpy()

    d = {}
    cases = [("white_2ch_22050", white((1, 2, 22050), 41)),
             ("guitar_2ch_22050", tr.cat([guitar(1, 22050, 42), guitar(1, 22050, 43) * 0.7], dim=1)),
             ("white_1ch_88200", white((1, 1, 88200), 44)),
             ("silence_and_click", tr.zeros((1, 2, 4096)))]
    cases[3][1][0, 1, 2000] = 0.9
    for i, (name, x) in enumerate(cases):
        ref = ref_logmel(x)
        fb_ref = net.spectrogram.mel_scale.fb.numpy()
        got32 = oracle.log_mel(x.numpy(), fb=fb_ref)
        got64 = oracle.log_mel(x.numpy(), fft_dtype=np.float64, fb=fb_ref)
        report(f"{name} (oracle fp32 FFT, own fb)", ref, oracle.log_mel(x.numpy()))
        report(f"{name} (oracle fp32 FFT)", ref, got32)
        report(f"{name} (oracle fp64 FFT)", ref, got64)
        d[f"x{i}"] = x.numpy()
        d[f"y{i}"] = ref
        d[f"name{i}"] = np.array(name)
    d["n"] = np.array(len(cases))
    fb_ref = net.spectrogram.mel_scale.fb.numpy()
    d["fb"] = fb_ref
    d["window"] = net.spectrogram.spectrogram.window.numpy()
    np.savez_compressed(os.path.join(HERE, "logmel.npz"), **d)
    fb = oracle.mel_filterbank()
    print(f"  mel fb: max|diff| = {np.abs(fb - fb_ref).max():.3e}, nnz ref={np.count_nonzero(fb_ref)} "
          f"oracle={np.count_nonzero(fb)}")
    win_ref = net.spectrogram.spectrogram.window.numpy()
    print(f"  hann: max|diff| = {np.abs(win_ref - oracle.hann_periodic(1024)).max():.3e}")


def lfo_estimates(B, n, seed, noise):
    """Signals shaped like the extractor's output: a 0.5-3 Hz cosine at the 172.5 Hz frame rate plus noise."""
    g = tr.Generator().manual_seed(seed)
    t = tr.arange(n) / 172.265625
    f = tr.exp(tr.rand(B, 1, generator=g) * math.log(6.0)) * 0.5
    ph = tr.rand(B, 1, generator=g) * 2 * math.pi
    amp = 0.3 + 0.2 * tr.rand(B, 1, generator=g)
    x = 0.5 + amp * tr.cos(2 * math.pi * f * t + ph) + noise * tr.randn(B, n, generator=g)
    return tr.clip(x, 0.0, 1.0).float()


def gen_logmel_c4():
    """Config-4-shaped rows (2 channels x 88200 samples -> (2, 256, 345)) of both audio families through the reference's
    own front end (models.py:170-175,199,207-208)."""
    print("M1 log-mel front end, config-4-shaped rows")
    net = rmodels.Spectral2DCNN(in_ch=2, n_samples=88200, sr=SR)
    net.eval()
    d = {}
    cases = [("white_2ch_88200", white((1, 2, 88200), 141)),
             ("guitar_2ch_88200", tr.cat([guitar(1, 88200, 142), guitar(1, 88200, 143) * 0.5], dim=1))]
    for i, (name, x) in enumerate(cases):
        with tr.no_grad():
            ref = tr.log(tr.clip(net.spectrogram(x), min=net.eps)).numpy()
        got64 = oracle.log_mel(x.numpy(), fft_dtype=np.float64, fb=net.spectrogram.mel_scale.fb.numpy())
        report(f"{name} (reference vs float64 FFT)", ref, got64)
        d[f"x{i}"] = x.numpy()
        d[f"y{i}"] = ref
        d[f"name{i}"] = np.array(name)
    d["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "logmel_c4.npz"), **d)


def gen_postproc():
    print("N4 smoothen / stretch_corners / find_valid_mod_sig_indices")
    d = {}
    k = 0
    # (noise, max_n_corners, smooth_n_frames): the shipped settings are (16, 0) with an 8- or 4-frame model
    # smoothing in front (configs/eval_em_unseen_effect.yml:95-100, eval_lfo.yml:126) and the defaults (10, 32)
    for noise, mx, sm in [(0.0, 10, 32), (0.0, 16, 0), (0.003, 16, 8), (0.01, 10, 32), (0.01, 16, 8), (0.0, 6, 4),
                          (0.05, 10, 32), (0.2, 16, 0)]:
        x = lfo_estimates(24, 345, 100 + k, noise)
        if k == 1:
            x[3] = 0.25                      # flat row: no corners, first == last
            x[4] = tr.linspace(0.1, 0.9, 345)  # ramp: no corners, one segment stretched onto itself
        ref = rmod.stretch_corners(x.clone(), max_n_corners=mx, smooth_n_frames=sm)
        got = oracle.stretch_corners(x.numpy(), mx, sm)
        report(f"stretch_corners noise={noise} max={mx} smooth={sm}", ref.numpy(), got)
        valid_in = rmod.find_valid_mod_sig_indices(rmod.smoothen(x, sm) if sm > 1 else x)
        valid_out = rmod.find_valid_mod_sig_indices(ref)
        assert valid_in == oracle.find_valid_mod_sig_indices(oracle.smoothen(x.numpy(), sm))
        assert valid_out == oracle.find_valid_mod_sig_indices(ref.numpy())
        print(f"    rows changed by the stretch: {int((ref != (rmod.smoothen(x, sm) if sm > 1 else x)).any(dim=1).sum())}/24, "
              f"valid before/after: {len(valid_in)}/{len(valid_out)}")
        d[f"x{k}"] = x.numpy()
        d[f"cfg{k}"] = np.array([mx, sm])
        d[f"y{k}"] = ref.numpy()
        d[f"valid_in{k}"] = np.array(valid_in, dtype=np.int64)
        d[f"valid_out{k}"] = np.array(valid_out, dtype=np.int64)
        k += 1
    d["n"] = np.array(k)
    x = tr.rand((6, 345), generator=tr.Generator().manual_seed(77))
    d["smooth_x"] = x.numpy()
    for w in (4, 8, 16, 32, 5, 12):
        ref = rmod.smoothen(x, w).numpy()
        report(f"smoothen window={w}", ref, oracle.smoothen(x.numpy(), w))
        d[f"smooth_y{w}"] = ref
    np.savez_compressed(os.path.join(HERE, "postproc.npz"), **d)


def gen_cnn():
    """N3: Spectral2DCNN (models.py:127-215) with seeded weights (tests/helpers.cnn_weights)."""
    print("N3 Spectral2DCNN body")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers

    def make(n_samples, n_mels, seed):
        net = rmodels.Spectral2DCNN(in_ch=2, n_samples=n_samples, sr=SR, n_mels=n_mels, kernel_size=(5, 13),
                                    out_channels=[64] * 6, temp_dilations=helpers.CNN_DILATIONS, pool_size=(2, 1),
                                    latent_dim=1, freq_mask_amount=0.25, time_mask_amount=0.25, use_ln=True)
        sd = helpers.cnn_weights(seed)
        missing = net.load_state_dict({k: tr.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not missing.unexpected_keys and all(k.startswith("spectrogram.") for k in missing.missing_keys)
        return net, sd

    d = {}
    # (a) reduced shape: 8192 samples, 64 mel bins -> 33 frames, final map 1 x 33
    net, sd = make(8192, 64, 7)
    net.eval()
    x = tr.cat([white((3, 1, 8192), 51), guitar(3, 8192, 52)], dim=1)
    with tr.no_grad():
        lm = tr.log(tr.clip(net.spectrogram(x), min=net.eps))
        y, lat = net(x)
    convs, ow, ob = helpers.cnn_oracle_args(sd)
    yo, lo = oracle.spectral_2dcnn_body(lm.numpy(), convs, ow, ob, helpers.CNN_DILATIONS)
    report("small: oracle fp32 output", y.numpy(), yo)
    report("small: oracle fp32 latent", lat.numpy(), lo)
    yo, lo = oracle.spectral_2dcnn_body(lm.numpy(), convs, ow, ob, helpers.CNN_DILATIONS, tf32_from_layer=1)
    report("small: oracle tf32 output", y.numpy(), yo)
    report("small: oracle tf32 latent", lat.numpy(), lo)
    d.update(small_x=x.numpy(), small_logmel=lm.numpy(), small_y=y.numpy(), small_latent=lat.numpy(),
             small_fb=net.spectrogram.mel_scale.fb.numpy())
    # (b) the same network in training mode: SpecAugment masks from the torch global generator (models.py:201-205)
    net.train()
    tr.manual_seed(1234)
    with tr.no_grad():
        y, lat = net(x)
    d.update(train_seed=np.array(1234), train_y=y.numpy(), train_latent=lat.numpy())
    # (c) the shipped shape: 88200 samples, 256 mel bins -> (B, 1, 345)
    net, sd = make(88200, 256, 8)
    net.eval()
    x = tr.cat([white((2, 1, 88200), 53), guitar(2, 88200, 54)], dim=1)
    with tr.no_grad():
        lm = tr.log(tr.clip(net.spectrogram(x), min=net.eps))
        y, lat = net(x)
    convs, ow, ob = helpers.cnn_oracle_args(sd)
    yo, lo = oracle.spectral_2dcnn_body(lm.numpy(), convs, ow, ob, helpers.CNN_DILATIONS)
    report("full: oracle fp32 output", y.numpy(), yo)
    report("full: oracle fp32 latent", lat.numpy(), lo)
    yo, lo = oracle.spectral_2dcnn_body(lm.numpy(), convs, ow, ob, helpers.CNN_DILATIONS, tf32_from_layer=1)
    report("full: oracle tf32 output", y.numpy(), yo)
    report("full: oracle tf32 latent", lat.numpy(), lo)
    d.update(full_x_seeds=np.array([53, 54]), full_x_probe=x.numpy()[..., ::4410], full_y=y.numpy(), full_latent=lat.numpy())
    np.savez_compressed(os.path.join(HERE, "cnn.npz"), **d)


if __name__ == "__main__":
    tr.manual_seed(0)
    oracle.build()
    which = sys.argv[1:] or ["lfo", "interp", "tremolo", "rng", "logmel", "logmel_c4", "fc", "postproc", "cnn"]
    fns = {"lfo": gen_lfo, "interp": gen_interp, "fc": gen_fc, "tremolo": gen_tremolo, "rng": gen_rng_lfos,
           "logmel": gen_logmel, "logmel_c4": gen_logmel_c4, "postproc": gen_postproc, "cnn": gen_cnn}
    for w in which:
        fns[w]()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"{f}: {os.path.getsize(os.path.join(HERE, f)) / 1024:.0f} KiB")
