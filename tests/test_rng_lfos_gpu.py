"""GPU parity of the RNG-driven LFO variants (quasi-periodic, combined) and corner finding.
The goldens hold the reference outputs under a fixed torch seed; the product draws from the same
torch global generator in the same order, so re-seeding must reproduce them."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import SHAPES6, golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_find_corners_matches_oracle():
    from mod_extraction_b200.modulations import find_corners
    rows = [oracle.make_mod_signal(882, 441.0, f, ph, s) for f, ph, s in
            [(2.0, 0.0, "tri"), (0.7, 1.0, "cos"), (2.9, 3.0, "saw"), (1.3, 5.0, "rsaw"), (1.1, 0.3, "sqr"),
             (2.2, 2.0, "rect_cos"), (0.6, 4.0, "inv_rect_cos")]]
    m = np.stack(rows)
    rt, rb = oracle.find_corners(m)
    t, b = find_corners(torch.from_numpy(m).to(DEV))
    assert np.array_equal(t.cpu().numpy(), rt) and np.array_equal(b.cpu().numpy(), rb)
    noise = np.random.RandomState(0).rand(4, 300).astype(np.float32)
    rt, rb = oracle.find_corners(noise)
    t, b = find_corners(torch.from_numpy(noise))           # CPU in -> CPU out
    assert not t.is_cuda and np.array_equal(t.numpy(), rt) and np.array_equal(b.numpy(), rb)


def test_quasi_periodic_reproduces_reference_under_seed():
    from mod_extraction_b200.modulations import make_quasi_periodic
    g = golden("rng_lfos")
    args = [float(v) for v in g["q_args"]]
    for k in range(int(g["q_n"])):
        torch.manual_seed(100 + k)                          # seed used by make_golden.py
        out = make_quasi_periodic(torch.from_numpy(g[f"q_base{k}"]).to(DEV), *args)
        assert np.array_equal(out.cpu().numpy(), g[f"q_out{k}"]), k


def test_quasi_periodic_batch_equals_sequential_calls():
    from mod_extraction_b200.modulations import make_quasi_periodic, make_quasi_periodic_batch
    g = golden("rng_lfos")
    args = [float(v) for v in g["q_args"]]
    base = torch.from_numpy(np.stack([g[f"q_base{k}"] for k in range(int(g["q_n"]))])).to(DEV)
    torch.manual_seed(7)
    seq = torch.stack([make_quasi_periodic(base[k], *args) for k in range(base.size(0))])
    torch.manual_seed(7)
    bat = make_quasi_periodic_batch(base, *args)
    assert torch.equal(seq, bat)
    flat = torch.full((2, 100), 0.5, device=DEV)            # no corners: returned unchanged
    assert torch.equal(make_quasi_periodic_batch(flat), flat)


def test_combined_reproduces_reference_under_seed():
    from mod_extraction_b200.modulations import make_combined_mod_sig
    g = golden("rng_lfos")
    for k in range(int(g["c_n"])):
        n, sr, f, ph = g[f"c_args{k}"]
        torch.manual_seed(200 + k)
        out = make_combined_mod_sig(int(n), float(sr), float(f), float(ph), SHAPES6)
        assert np.abs(out.cpu().numpy() - g[f"c_out{k}"]).max() <= 1e-6, k     # LFO tolerance (cos ulp)


def test_combined_batch_equals_sequential_calls():
    from mod_extraction_b200.modulations import make_combined_mod_sig, make_combined_mod_sig_batch
    rng = np.random.RandomState(1)
    f = np.exp(rng.uniform(np.log(1.0), np.log(3.0), 16))
    ph = rng.uniform(0, 2 * np.pi, 16)
    torch.manual_seed(11)
    seq = torch.stack([make_combined_mod_sig(882, 441, f[i], ph[i], SHAPES6) for i in range(16)])
    torch.manual_seed(11)
    bat = make_combined_mod_sig_batch(882, 441, f, ph, SHAPES6)
    assert torch.equal(seq, bat)
    assert float(bat.min()) >= 0.0 and float(bat.max()) <= 1.0


@pytest.mark.parametrize("B,seed", [(1, 3), (64, 5), (4096, 43)])
def test_combined_device_replay_equals_host_replay(B, seed):
    """The device replay of make_combined_mod_sig's draw order (raw generator words handed to the GPU) gives the signals,
    the base shapes and the final generator state of the per-example python loop of scalar torch.randint calls."""
    from mod_extraction_b200.modulations import make_combined_mod_sig_batch
    rng = np.random.RandomState(seed)
    f = np.exp(rng.uniform(np.log(1.0), np.log(3.0), B))
    ph = rng.uniform(0, 2 * np.pi, B)
    torch.manual_seed(seed)
    torch.rand(seed * 37)                                       # start somewhere inside a generator block
    slow, base_slow = make_combined_mod_sig_batch(882, 441, f, ph, SHAPES6, return_base=True, host_replay=True)
    next_slow = torch.rand(5)
    torch.manual_seed(seed)
    torch.rand(seed * 37)
    fast, base_fast = make_combined_mod_sig_batch(882, 441, f, ph, SHAPES6, return_base=True)
    next_fast = torch.rand(5)
    assert torch.equal(base_slow.cpu(), base_fast.cpu())
    assert torch.equal(slow, fast)
    assert torch.equal(next_slow, next_fast), "the generator must end where the reference's loop leaves it"


def test_combined_deferred_read_back_equals_blocking_call():
    """deferred=True hands the signals back before the 8-byte read-back; finish() then leaves the generator where the
    blocking call does."""
    from mod_extraction_b200.modulations import make_combined_mod_sig_batch
    rng = np.random.RandomState(8)
    f = np.exp(rng.uniform(np.log(1.0), np.log(3.0), 300))
    ph = rng.uniform(0, 2 * np.pi, 300)
    torch.manual_seed(21)
    ref = make_combined_mod_sig_batch(882, 441, f, ph, SHAPES6)
    next_ref = torch.rand(5)
    torch.manual_seed(21)
    out, finish = make_combined_mod_sig_batch(882, 441, f, ph, SHAPES6, deferred=True)
    doubled = out * 2                                           # work queued before the read-back
    assert finish() is True
    assert torch.equal(out, ref) and torch.equal(doubled, ref * 2)
    assert torch.equal(torch.rand(5), next_ref)
