"""Bulk access to torch's global CPU generator, word for word.

"Identical seeds" in the reference means the torch *global CPU* generator (util.py:38-49), drawn from one scalar
at a time inside python loops (``util.randint`` / ``util.sample_uniform`` with n = 1).  That generator is a plain
MT19937: every ``torch.rand(1)`` (float32) and every ``torch.randint(lo, hi, (1,))`` with ``hi - lo < 2**28`` consumes
exactly ONE 32-bit output word -- ``(w & 0xFFFFFF) * 2**-24`` and ``w % (hi - lo) + lo`` respectively (ATen
``uniform_real_distribution<float>`` / ``uniform_int_from_to_distribution``; pinned by tests/test_rng_cpu.py).

``TorchMT`` reads the generator state, produces the next ``n`` raw words with numpy's MT19937 in one call (the same
stream the scalar calls would see), and ``consume(n)`` writes the advanced state back, so vectorised host code and
device kernels that were handed the raw words leave the generator exactly where the reference's python loop would.

State layout (at::CPUGeneratorImplState, 5056 bytes): seed u64 @0, left i32 @8, seeded i32 @12, next u64 @16,
state[624] as u64 @24; at::mt19937 keeps ``left + next == 625`` inside a block and regenerates when ``left == 1``.
"""
from __future__ import annotations

import numpy as np
import torch

_N = 624
_OFF_LEFT, _OFF_NEXT, _OFF_KEY = 8, 16, 24


class TorchMT:
    def __init__(self) -> None:
        self._state = torch.get_rng_state().numpy().copy()
        if self._state.size != 5056:                    # another torch build: refuse rather than guess
            raise RuntimeError(f"unexpected CPU generator state size {self._state.size}")
        left = int(self._state[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0])
        nxt = int(self._state[_OFF_NEXT:_OFF_NEXT + 8].view(np.uint64)[0])
        key = self._state[_OFF_KEY:_OFF_KEY + _N * 8].view(np.uint64).astype(np.uint32)
        self._bg = np.random.MT19937()
        s = self._bg.state
        s["state"]["key"] = key
        s["state"]["pos"] = _N if left == 1 else nxt
        self._bg.state = s
        self._start = self._bg.state
        self._drawn = 0

    def words(self, n: int) -> np.ndarray:
        """The next ``n`` raw 32-bit words after everything handed out so far (uint32)."""
        self._drawn += n
        return self._bg.random_raw(n).astype(np.uint32)

    def consume(self, n: int) -> None:
        """Advance torch's generator by ``n`` words (counted from where it stood at construction)."""
        if n == 0:
            return
        bg = np.random.MT19937()
        bg.state = self._start
        bg.random_raw(n)
        s = bg.state["state"]
        pos = int(s["pos"])
        st = self._state
        st[_OFF_KEY:_OFF_KEY + _N * 8].view(np.uint64)[:] = s["key"].astype(np.uint64)
        st[_OFF_NEXT:_OFF_NEXT + 8].view(np.uint64)[0] = pos
        st[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0] = _N + 1 - pos
        torch.set_rng_state(torch.from_numpy(st.copy()))


def words_to_uniform(w: np.ndarray, low: float, high: float) -> np.ndarray:
    """What ``util.sample_uniform(low, high)`` returns for each word: float32 arithmetic, then a python float."""
    x = (w & np.uint32(0xFFFFFF)).astype(np.float32) * np.float32(2.0 ** -24)
    return (x * np.float32(high - low) + np.float32(low)).astype(np.float64)


def words_to_randint(w: np.ndarray, low: int, high: int) -> np.ndarray:
    """What ``util.randint(low, high)`` returns for each word (``high - low`` < 2**28)."""
    assert 0 < high - low < (1 << 28)
    return (w % np.uint32(high - low)).astype(np.int64) + low
