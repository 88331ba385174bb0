"""Shared test helpers: golden loading, synthetic audio families, error metrics."""
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SR = 44100
SHAPES6 = ["cos", "tri", "rect_cos", "inv_rect_cos", "saw", "rsaw"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def white(shape, seed):
    """Family W: white noise in [-0.5, 0.5)."""
    rng = np.random.RandomState(seed)
    return ((rng.random_sample(shape) * 2.0 - 1.0) * 0.5).astype(np.float32)


def guitar(B, N, seed):
    """Family G: decaying harmonic partials + -60 dB noise."""
    rng = np.random.RandomState(seed)
    t = np.arange(N, dtype=np.float64) / SR
    out = np.empty((B, 1, N), dtype=np.float32)
    for b in range(B):
        f0 = 110.0 * (1.0 + 0.25 * (b % 8))
        sig = sum((0.3 / k) * np.sin(2 * math.pi * f0 * k * t) * np.exp(-k * t) for k in range(1, 9))
        out[b, 0] = (sig + 1e-3 * rng.standard_normal(N)).astype(np.float32)
    return out


def snr_db(ref, got):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(got, dtype=np.float64) - ref
    den = float((err ** 2).sum())
    if den == 0.0:
        return float("inf")
    return 10.0 * math.log10(float((ref ** 2).sum()) / den)


def fc_params_from_golden(g, k):
    """Effect parameters of flanger_chorus.npz case k: 0-d arrays were python floats."""
    params = []
    j = 0
    while f"p{k}_{j}" in g.files:
        p = g[f"p{k}_{j}"]
        params.append(float(p) if p.ndim == 0 else p.astype(np.float32))
        j += 1
    return params


CNN_DILATIONS = [1, 1, 2, 4, 8, 16]       # configs/models/spectral_2dcnn.yml


def cnn_weights(seed, in_ch=2, n_layers=6, ch=64, kh=5, kw=13, latent_dim=1):
    """Seeded random weights for Spectral2DCNN as a state-dict of numpy arrays (reference key names,
    models.py:183-195): lively enough that every layer matters (fan-in scaled convs, PReLU slopes 0.05..0.4)."""
    rng = np.random.RandomState(seed)
    sd = {}
    c_in = in_ch
    for i in range(n_layers):
        bound = 1.7 / math.sqrt(c_in * kh * kw)
        sd[f"cnn.{4 * i + 1}.weight"] = rng.uniform(-bound, bound, (ch, c_in, kh, kw)).astype(np.float32)
        sd[f"cnn.{4 * i + 1}.bias"] = rng.uniform(-0.1, 0.1, (ch,)).astype(np.float32)
        sd[f"cnn.{4 * i + 3}.weight"] = rng.uniform(0.05, 0.4, (ch,)).astype(np.float32)
        c_in = ch
    sd["output.weight"] = rng.uniform(-0.5, 0.5, (latent_dim, ch, 1)).astype(np.float32)
    sd["output.bias"] = rng.uniform(-0.1, 0.1, (latent_dim,)).astype(np.float32)
    return sd


def cnn_oracle_args(sd, n_layers=6):
    """cnn_weights() -> the (convs, out_w, out_b) arguments of oracle.spectral_2dcnn_body."""
    convs = [(sd[f"cnn.{4 * i + 1}.weight"], sd[f"cnn.{4 * i + 1}.bias"], sd[f"cnn.{4 * i + 3}.weight"])
             for i in range(n_layers)]
    return convs, sd["output.weight"][:, :, 0], sd["output.bias"]
