"""Host logic: bulk access to torch's global CPU generator (mod_extraction_b200/_rng.py) yields the very words the
reference's scalar util.randint / util.sample_uniform calls consume (util.py:38-49), and leaves the generator where
they would."""
import math

import numpy as np
import pytest
import torch

from mod_extraction_b200 import util
from mod_extraction_b200._rng import TorchMT, words_to_randint, words_to_uniform


@pytest.mark.parametrize("seed,burn", [(0, 0), (43, 1), (7, 623), (7, 624), (7, 625), (123, 5000)])
def test_words_are_what_scalar_draws_consume(seed, burn):
    torch.manual_seed(seed)
    if burn:
        torch.rand(burn)
    state = torch.get_rng_state()
    mt = TorchMT()
    w = mt.words(1500)
    # interleaved scalar draws exactly like the reference's loops: uniform, randint, uniform, ...
    got_u, got_i = [], []
    for k in range(500):
        got_u.append(util.sample_uniform(0.0, 2 * math.pi))
        got_i.append(util.randint(0, 6))
        got_u.append(util.sample_uniform(0.5, 3.0))
    exp_u = np.empty(1000)
    exp_u[0::2] = words_to_uniform(w[0::3], 0.0, 2 * math.pi)
    exp_u[1::2] = words_to_uniform(w[2::3], 0.5, 3.0)
    assert np.array_equal(np.array(got_u), exp_u)
    assert np.array_equal(np.array(got_i), words_to_randint(w[1::3], 0, 6))
    after_scalar = torch.get_rng_state()
    # consume() moves the generator to the same place
    torch.set_rng_state(state)
    mt2 = TorchMT()
    mt2.words(1500)
    mt2.consume(1500)
    assert torch.equal(torch.get_rng_state()[8:24 + 624 * 8], after_scalar[8:24 + 624 * 8])
    assert torch.equal(torch.rand(700), (torch.set_rng_state(after_scalar), torch.rand(700))[1])


def test_consume_partial_and_zero():
    torch.manual_seed(99)
    mt = TorchMT()
    w = mt.words(2000)
    mt.consume(0)
    assert int(torch.randint(0, 1 << 20, (1,))) == int(w[0] % (1 << 20))     # nothing consumed by consume(0)
    torch.manual_seed(99)
    mt = TorchMT()
    w = mt.words(2000)                      # handing out more words than are consumed is fine
    mt.consume(777)
    assert int(torch.randint(0, 1 << 20, (1,))) == int(w[777] % (1 << 20))
