"""GPU parity of the extracted-LFO post-processing (smoothen, stretch_corners,
find_valid_mod_sig_indices; reference modulations.py:259-362) against the reference goldens and the oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def lfo_estimates(B, n, seed, noise):
    rs = np.random.RandomState(seed)
    t = np.arange(n) / 172.265625
    f = np.exp(rs.rand(B, 1) * math.log(6.0)) * 0.5
    ph = rs.rand(B, 1) * 2 * math.pi
    amp = 0.3 + 0.2 * rs.rand(B, 1)
    x = 0.5 + amp * np.cos(2 * math.pi * f * t + ph) + noise * rs.randn(B, n)
    return np.clip(x, 0.0, 1.0).astype(np.float32)


def test_smoothen_matches_reference_goldens():
    from mod_extraction_b200.modulations import smoothen
    g = golden("postproc")
    x = torch.from_numpy(g["smooth_x"]).to(DEV)
    for w in (4, 8, 16, 32):
        assert np.array_equal(smoothen(x, w).cpu().numpy(), g[f"smooth_y{w}"]), w
    for w in (5, 12):
        assert np.abs(smoothen(x, w).cpu().numpy() - g[f"smooth_y{w}"]).max() <= 2e-7, w
    assert smoothen(x, 1) is x and smoothen(x, 0) is x                      # modulations.py:359
    y = smoothen(torch.from_numpy(g["smooth_x"]), 8)                        # CPU in -> CPU out
    assert not y.is_cuda and np.array_equal(y.numpy(), g["smooth_y8"])
    full = smoothen(x, x.size(1))                                            # window == n: one mean per row
    assert full.shape == (x.size(0), 1)
    assert np.array_equal(full.cpu().numpy(), oracle.smoothen(g["smooth_x"], x.size(1)))
    x3 = torch.from_numpy(g["smooth_x"]).to(DEV).reshape(2, 3, -1)          # leading dims kept
    assert np.array_equal(smoothen(x3, 8).cpu().numpy(), g["smooth_y8"].reshape(2, 3, -1))


def test_stretch_corners_matches_reference_goldens():
    from mod_extraction_b200.modulations import stretch_corners
    g = golden("postproc")
    for k in range(int(g["n"])):
        mx, sm = (int(v) for v in g[f"cfg{k}"])
        out = stretch_corners(torch.from_numpy(g[f"x{k}"]).to(DEV), max_n_corners=mx, smooth_n_frames=sm)
        assert np.array_equal(out.cpu().numpy(), g[f"y{k}"], equal_nan=True), k


def test_find_valid_indices_match_reference_goldens():
    from mod_extraction_b200.modulations import find_valid_mod_sig_indices, smoothen
    g = golden("postproc")
    for k in range(int(g["n"])):
        sm = int(g[f"cfg{k}"][1])
        x = torch.from_numpy(g[f"x{k}"]).to(DEV)
        assert find_valid_mod_sig_indices(smoothen(x, sm)) == g[f"valid_in{k}"].tolist(), k
        assert find_valid_mod_sig_indices(torch.from_numpy(g[f"y{k}"])) == g[f"valid_out{k}"].tolist(), k


@pytest.mark.parametrize("noise,mx,sm", [(0.0, 16, 0), (0.005, 16, 8), (0.02, 10, 32), (0.3, 62, 0)])
def test_large_batch_matches_oracle(noise, mx, sm):
    from mod_extraction_b200.modulations import find_valid_mod_sig_indices, stretch_corners
    x = lfo_estimates(1024, 345, 7, noise)
    out = stretch_corners(torch.from_numpy(x).to(DEV), max_n_corners=mx, smooth_n_frames=sm)
    ref = oracle.stretch_corners(x, mx, sm)
    assert np.array_equal(out.cpu().numpy(), ref, equal_nan=True)
    assert find_valid_mod_sig_indices(out) == oracle.find_valid_mod_sig_indices(ref)


def test_edge_shapes():
    from mod_extraction_b200.modulations import find_valid_mod_sig_indices, mod_sig_to_corners, stretch_corners
    rs = np.random.RandomState(3)
    for n in (2, 3, 4, 33, 64, 65, 1000):
        x = rs.rand(5, n).astype(np.float32)
        out = stretch_corners(torch.from_numpy(x).to(DEV), max_n_corners=62, smooth_n_frames=0)
        assert np.array_equal(out.cpu().numpy(), oracle.stretch_corners(x, 62, 0), equal_nan=True), n
        assert find_valid_mod_sig_indices(torch.from_numpy(x).to(DEV)) == oracle.find_valid_mod_sig_indices(x), n
    empty = stretch_corners(torch.zeros((0, 345), device=DEV), 10, 32)
    assert empty.shape == (0, 314)
    with pytest.raises(RuntimeError):
        stretch_corners(torch.zeros((2, 345), device=DEV), max_n_corners=1000, smooth_n_frames=0)
    m = torch.from_numpy(oracle.make_mod_signal(882, 441.0, 1.5, 0.3, "tri")[None]).to(DEV)
    t, b = mod_sig_to_corners(m, 345)
    rt, rb = oracle.find_corners(oracle.linear_interpolate_last_dim(m.cpu().numpy(), 345))
    assert np.array_equal(t.cpu().numpy(), rt) and np.array_equal(b.cpu().numpy(), rb)
