"""GPU phaser vs this repo's own CPU restatement of pedalboard/JUCE (PARITY UNPINNED):
the chunked affine scan re-associates the recurrence, so the bar is the north_star tolerance
(1e-4 max-abs, SNR >= 80 dB), not bit-exactness."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import guitar, snr_db, white

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _params(B, seed):
    rng = np.random.RandomState(seed)
    logu = lambda lo, hi: np.exp(rng.uniform(np.log(lo), np.log(hi), B)).astype(np.float32)
    U = lambda lo, hi: rng.uniform(lo, hi, B).astype(np.float32)
    # configs/train_lfo_phaser.yml:33-48
    return logu(0.5, 3.0), U(0.2, 1.0), logu(70.0, 18000.0), U(0.0, 0.7), U(0.2, 1.0)


@pytest.mark.parametrize("family", ["white", "guitar"])
@pytest.mark.parametrize("N", [88200, 88200 + 14700, 5000, 127, 129])
def test_phaser_vs_own_oracle(family, N):
    from mod_extraction_b200.phaser import Phaser
    B = 16
    x = white((B, N), 3) if family == "white" else guitar(B, N, 4)[:, 0]
    ps = _params(B, N)
    ref = oracle.phaser(x, 44100.0, *ps)
    y = Phaser(44100.0)(torch.from_numpy(x).to(DEV), *[torch.from_numpy(p) for p in ps]).cpu().numpy()
    err = np.abs(y - ref).max()
    assert err <= 1e-4 and snr_db(ref, y) >= 80.0, (family, N, float(err), snr_db(ref, y))


def test_phaser_extreme_parameters():
    from mod_extraction_b200.phaser import Phaser
    N = 30000
    x = white((6, N), 9)
    rate = np.array([0.5, 3.0, 3.0, 0.5, 1.0, 2.0], dtype=np.float32)
    depth = np.array([1.0, 1.0, 0.2, 0.2, 1.0, 0.0], dtype=np.float32)
    centre = np.array([70.0, 18000.0, 20.0, 1000.0, 300.0, 5000.0], dtype=np.float32)
    fb = np.array([0.7, 0.7, 0.0, 0.69, 0.7, 0.3], dtype=np.float32)
    mix = np.array([1.0, 0.2, 1.0, 0.5, 0.0, 1.0], dtype=np.float32)
    ref = oracle.phaser(x, 44100.0, rate, depth, centre, fb, mix)
    y = Phaser(44100.0)(torch.from_numpy(x).to(DEV), rate, depth, centre, fb, mix).cpu().numpy()
    assert np.abs(y - ref).max() <= 1e-4
    assert np.array_equal(y[4], np.clip(x[4], -1, 1))       # mix = 0 -> dry, exactly


def test_phaser_subset_and_shapes():
    from mod_extraction_b200.phaser import Phaser
    B, N = 8, 20000
    x = torch.from_numpy(white((B, 1, N), 5)).to(DEV)
    ps = [torch.from_numpy(p) for p in _params(B, 1)]
    ph = Phaser(44100.0)
    y = ph(x, *ps)
    assert y.shape == (B, 1, N)
    out = torch.zeros_like(x)
    ph(x, *ps, example_index=torch.tensor([1, 6]), out=out)
    assert torch.equal(out[1], y[1]) and torch.equal(out[6], y[6]) and float(out[0].abs().max()) == 0.0
    ycpu = ph(x.cpu(), *ps)
    assert not ycpu.is_cuda and torch.equal(ycpu, y.cpu())


def test_apply_pedalboard_phaser_draw_order():
    """Same host RNG draws, in the order of datasets.py:461-465."""
    from mod_extraction_b200.phaser import apply_pedalboard_phaser
    ranges = {"depth": {"min": 0.2, "max": 1.0}, "centre_frequency_hz": {"min": 70.0, "max": 18000.0},
              "feedback": {"min": 0.0, "max": 0.7}, "mix": {"min": 0.2, "max": 1.0}}
    x = torch.from_numpy(white((1, 30000), 8))
    torch.manual_seed(5)
    np.random.seed(5)
    y, p = apply_pedalboard_phaser(x, 44100.0, 1.7, ranges)
    torch.manual_seed(5)
    np.random.seed(5)
    depth = (torch.rand(1) * (1.0 - 0.2) + 0.2).item()
    from scipy.stats import loguniform
    centre = float(loguniform.rvs(70.0, 18000.0, size=1)[0])
    feedback = (torch.rand(1) * (0.7 - 0.0) + 0.0).item()
    mix = (torch.rand(1) * (1.0 - 0.2) + 0.2).item()
    assert p == {"depth": depth, "feedback": feedback, "mix": mix, "rate_hz": 1.7, "shape": "cos"}
    ref = oracle.phaser(x.numpy(), 44100.0, 1.7, depth, centre, feedback, mix)
    assert np.abs(y.numpy() - ref).max() <= 1e-4
