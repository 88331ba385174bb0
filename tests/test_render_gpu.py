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
    ref_wet, ref_lm = bench.oracle_step(dry, effect, mod_lo, fc, ph, threads=4)
    fcx = np.nonzero(effect != 2)[0]
    assert np.array_equal(wa.cpu().numpy()[fcx], ref_wet[fcx])
    err = np.abs(la.cpu().numpy() - ref_lm)
    assert (err <= 1e-4).mean() >= 0.995
