"""GPU: rows N1 / N2 -- batched parameter sampling with the reference's draw order, and the cache format."""
import hashlib
import json
import math

import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import SHAPES6, SR, white

pytestmark = pytest.mark.gpu

FLANGER_CFG = {"max_min_delay_ms": 1.0, "max_lfo_delay_ms": 10.0,          # configs/data/gen_idmt_fl.yml:34-51
               "feedback": {"min": 0.0, "max": 0.7}, "min_delay_width": {"min": 0.0, "max": 1.0},
               "width": {"min": 0.25, "max": 1.0}, "depth": {"min": 0.25, "max": 1.0}, "mix": {"min": 0.25, "max": 1.0}}
MOD_CFG = {"rate_hz": {"min": 0.5, "max": 3.0}, "phase": {"min": 0.0, "max": 6.28318530718}, "shapes": SHAPES6,
           "exp": 1.0}


def test_flanger_render_step_matches_reference_draw_order_and_oracle():
    from mod_extraction_b200.data import FlangerRenderStep, sample_mod_sig_batch
    B, N = 6, 8800
    torch.manual_seed(3)
    np.random.seed(3)
    mod_lo, fxp = sample_mod_sig_batch(MOD_CFG, B, N, SR)
    assert mod_lo.shape == (B, N // 100) and fxp["rate_hz"].dtype == torch.float64 and len(fxp["shape"]) == B
    dry = torch.from_numpy(white((B, 1, N), 1))
    torch.manual_seed(9)
    step = FlangerRenderStep({"flanger": FLANGER_CFG}, B, N, SR)
    d2, wet, m2, fx = step((dry, mod_lo, fxp))
    # data_modules.py:421-445: five sample_uniform(n=B) draws in this order
    torch.manual_seed(9)
    exp = {}
    for k in ("feedback", "min_delay_width", "width", "depth", "mix"):
        lo, hi = FLANGER_CFG[k]["min"], FLANGER_CFG[k]["max"]
        exp[k] = torch.rand(B) * (hi - lo) + lo
        assert torch.equal(fx[k], exp[k]), k
    assert set(fx) >= {"rate_hz", "phase", "shape", "exp", "depth", "feedback", "max_lfo_delay_ms", "max_min_delay_ms",
                       "min_delay_width", "mix", "width"}
    ref = oracle.flanger_chorus(dry.numpy(), oracle.linear_interpolate_last_dim(mod_lo.cpu().numpy(), N),
                                *[exp[k].numpy() for k in ("feedback", "min_delay_width", "width", "depth", "mix")],
                                max_min_delay_ms=1.0, max_lfo_delay_ms=10.0)
    assert not wet.is_cuda and np.array_equal(wet.numpy(), ref)


def test_sample_mod_sig_draws_like_reference_getitem():
    """datasets.py:367-372: rate (numpy/scipy RNG), phase and shape (torch RNG) per example."""
    from mod_extraction_b200.data import sample_mod_sig_batch
    from scipy.stats import loguniform
    B = 4
    torch.manual_seed(1)
    np.random.seed(1)
    mod, fxp = sample_mod_sig_batch(MOD_CFG, B, 88200, SR)
    torch.manual_seed(1)
    np.random.seed(1)
    for b in range(B):
        rate = float(loguniform.rvs(0.5, 3.0, size=1)[0])
        phase = (torch.rand(1) * (6.28318530718 - 0.0) + 0.0).item()
        shape = SHAPES6[torch.randint(0, 6, (1,)).item()]
        assert fxp["rate_hz"][b].item() == rate and fxp["phase"][b].item() == phase and fxp["shape"][b] == shape
        ref = oracle.make_mod_signal(882, 441, rate, phase, shape)
        assert np.abs(mod[b].cpu().numpy() - ref).max() <= 1e-6


def test_cache_roundtrip_and_naming(tmp_path):
    from mod_extraction_b200.data import PreprocessedDataset, write_cache
    B, N = 3, 4400
    dry = torch.from_numpy(white((B, 1, N), 4))
    wet = torch.from_numpy(white((B, 1, N), 5))
    mod = torch.rand(B, N // 100)
    fx = {"rate_hz": torch.tensor([1.0, 2.0, 0.7], dtype=torch.float64), "phase": torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64),
          "shape": ["cos", "tri", "saw"], "exp": 1.0, "depth": torch.tensor([0.5, 0.6, 0.7]), "max_lfo_delay_ms": 10.0}
    stems = write_cache(str(tmp_path), dry, wet, mod, fx, SR)
    # scripts/scratch.py:151-158 naming
    f0 = {"rate_hz": 1.0, "phase": 0.1, "shape": "cos", "exp": 1.0, "depth": fx["depth"][0].item(), "max_lfo_delay_ms": 10.0}
    want = hashlib.md5(json.dumps({k: str(v) for k, v in f0.items()}, sort_keys=True).encode("utf-8")).hexdigest()
    assert stems[0] == want
    ds = PreprocessedDataset(str(tmp_path), N, SR)
    assert len(ds) == B
    seen = 0
    for i in range(B):
        d, w, m, f = ds[i]
        j = stems.index(ds.pt_paths[i].split("/")[-1][:-3])
        assert torch.equal(d, dry[j]) and torch.equal(w, wet[j]) and torch.equal(m, mod[j])
        assert f["shape"] == fx["shape"][j] and f["exp"] == 1.0
        seen += 1
    assert seen == B


class _LiveDraws:
    """The reference's util.sample_uniform / util.choice on the torch global CPU generator (util.py:32-49)."""

    def uniform(self, low, high):
        return ((torch.rand(1) * (high - low)) + low).item()

    def choice(self, n):
        return torch.randint(low=0, high=n, size=(1,)).item()


@pytest.mark.parametrize("kind", ["combined", "quasiperiodic"])
def test_sample_mod_sig_follows_the_reference_stream_example_by_example(kind):
    """datasets.py:365-398 for the eval_lfo_combined / eval_lfo_quasi configurations: every example's triple AND its
    section draws are contiguous in the torch stream.  The restatement below walks the stream like __getitem__ does."""
    from mod_extraction_b200.data import sample_mod_sig_batch
    from scipy.stats import loguniform
    cfg = dict(MOD_CFG)
    cfg["rate_hz"] = {"min": 1.0, "max": 3.0} if kind == "combined" else {"min": 0.5, "max": 2.0}
    q = dict(l_min=0.10, l_max=0.3333, r_min=0.10, r_max=0.3333, lr_split=0.5)      # configs/eval_lfo_quasi.yml:49-54
    cfg[kind] = True
    cfg.update(q)
    B, N = 12, 88200
    torch.manual_seed(21)
    np.random.seed(21)
    mod, fxp = sample_mod_sig_batch(cfg, B, N, SR)
    tail = torch.rand(3)
    torch.manual_seed(21)
    np.random.seed(21)
    draws = _LiveDraws()
    for b in range(B):
        rate = float(loguniform.rvs(cfg["rate_hz"]["min"], cfg["rate_hz"]["max"], size=1)[0])
        phase = draws.uniform(0.0, 6.28318530718)
        shape = SHAPES6[draws.choice(6)]
        assert fxp["rate_hz"][b].item() == rate and fxp["phase"][b].item() == phase and fxp["shape"][b] == shape, b
        if kind == "combined":
            ref = oracle.make_combined_mod_sig(882, 441, rate, phase, SHAPES6, rng=draws)
        else:
            ref = oracle.make_quasi_periodic(oracle.make_mod_signal(882, 441, rate, phase, shape), rng=draws, **q)
        assert np.abs(mod[b].cpu().numpy() - ref).max() <= 1e-6, b
    assert torch.equal(tail, torch.rand(3)), "generator state after the batch"
    # the one-launch form draws the same triples only for batch 1 (documented)
    torch.manual_seed(21)
    np.random.seed(21)
    mod1, _ = sample_mod_sig_batch(cfg, 1, N, SR, exact_stream=False)
    assert torch.equal(mod1[0], mod[0])
