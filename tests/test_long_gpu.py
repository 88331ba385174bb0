"""GPU: BASELINE config 5 shape -- 60 s clips (2 646 000 samples, 26 460 control points) through every kernel."""
import math

import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import SHAPES6, SR, golden, snr_db, white

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N60 = 60 * SR


def test_flanger_and_chorus_60s_bit_exact():
    from mod_extraction_b200.fx import MonoFlangerChorusModule
    rng = np.random.RandomState(60)
    B = 3
    x = white((B, 1, N60), 60)
    lo = np.stack([oracle.make_mod_signal(N60 // 100, SR / 100, f, ph, s)
                   for f, ph, s in [(0.5, 1.0, "tri"), (2.9, 4.0, "saw"), (1.3, 0.2, "rsaw")]])
    U = lambda a, b: rng.uniform(a, b, B).astype(np.float32)
    for mmd, mld, mdw_lo in [(1.0, 10.0, 0.0), (30.0, 10.0, 0.367)]:
        params = [U(0.0, 0.7), U(mdw_lo, 1.0), U(0.25, 1.0), U(0.25, 1.0), U(0.25, 1.0)]
        ref = oracle.flanger_chorus(x, oracle.linear_interpolate_last_dim(lo, N60), *params, max_min_delay_ms=mmd,
                                    max_lfo_delay_ms=mld)
        m = MonoFlangerChorusModule(B, 1, N60, SR, mmd, mld)
        y = m.forward_control_rate(torch.from_numpy(x).to(DEV), torch.from_numpy(lo).to(DEV),
                                   *[torch.from_numpy(p).to(DEV) for p in params]).cpu().numpy()
        assert np.array_equal(y, ref), (mmd, mld)


def test_phaser_60s_vs_own_oracle():
    from mod_extraction_b200.phaser import Phaser
    B = 3
    x = white((B, N60), 61)
    rate = np.array([0.5, 1.7, 3.0], dtype=np.float32)
    depth = np.array([1.0, 0.5, 0.2], dtype=np.float32)
    centre = np.array([70.0, 1300.0, 18000.0], dtype=np.float32)
    fb = np.array([0.7, 0.3, 0.0], dtype=np.float32)
    mix = np.array([1.0, 0.6, 0.2], dtype=np.float32)
    ref = oracle.phaser(x, float(SR), rate, depth, centre, fb, mix)
    y = Phaser(float(SR))(torch.from_numpy(x).to(DEV), rate, depth, centre, fb, mix).cpu().numpy()
    err = np.abs(y - ref).max()
    assert err <= 1e-4 and snr_db(ref, y) >= 80.0, (float(err), snr_db(ref, y))


def test_logmel_60s_shape_and_sampled_frames():
    from mod_extraction_b200.models import LogMelSpectrogram
    g = golden("logmel")
    front = LogMelSpectrogram(fb=torch.from_numpy(g["fb"]), window=torch.from_numpy(g["window"])).to(DEV)
    x = white((2, 1, N60), 62)
    y = front(torch.from_numpy(x).to(DEV))
    assert y.shape == (2, 1, 256, N60 // 256 + 1)            # 10 336 frames
    # frames away from the clip edges depend only on their own 1024 samples: compare a window of them
    # with the oracle run on the corresponding excerpt
    t0, nt = 5000, 40
    seg = x[:, :, t0 * 256 - 512 - 2048: (t0 + nt) * 256 + 512 + 2048]
    ref = oracle.log_mel(seg, fb=g["fb"], fft_dtype=np.float64)
    off = (512 + 2048) // 256                             # frame t of the clip is frame t - t0 + off of the excerpt
    got = y[:, :, :, t0:t0 + nt].cpu().numpy()
    err = np.abs(got - ref[:, :, :, off:off + nt])
    assert (err <= 1e-4).mean() >= 0.9999 and err.max() <= 2e-3        # white noise, vs float64 FFT; see test_logmel_gpu
