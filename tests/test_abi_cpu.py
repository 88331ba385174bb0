"""CPU tests: the C-ABI library loads and exports every symbol include/modfx.h declares."""
import os
import re

import pytest

from mod_extraction_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "modfx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(modfx_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    _build.build()
    L = _lib.lib()
    names = header_symbols()
    assert "modfx_flanger_chorus_f32" in names and "modfx_logmel_f32" in names
    for n in names:
        assert hasattr(L, n), f"libmodfx.so does not export {n}"
    assert sorted(_lib.exported_symbols()) == names, "ctypes signature table and header disagree"
    assert L.modfx_abi_version() == 1


def test_no_cpu_fallback():
    """Without a GPU the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mod_extraction_b200 import fx, modulations, util
    m = fx.MonoFlangerChorusModule(1, 1, 64, 44100, 1.0, 10.0)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 64), torch.zeros(1, 64))
    with pytest.raises(RuntimeError):
        modulations.make_mod_signal(10, 441, 2.0)
    with pytest.raises(RuntimeError):
        util.linear_interpolate_last_dim(torch.zeros(2, 8), 16)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mod_extraction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_reference_asserts_are_kept():
    import torch
    from mod_extraction_b200 import fx
    m = fx.MonoFlangerChorusModule(2, 1, 64, 44100, 1.0, 10.0)
    x = torch.zeros(2, 1, 64)
    with pytest.raises(AssertionError):
        m(x, torch.zeros(2, 63))                       # fx.py:83
    with pytest.raises(AssertionError):
        m(x, torch.zeros(2, 64), feedback=1.0)         # fx.py:69 (feedback < 1 strictly)
    with pytest.raises(AssertionError):
        m(x, torch.zeros(2, 64), mix=torch.tensor([0.5, 1.5]))     # fx.py:55
    with pytest.raises(AssertionError):
        m(x[:, 0], torch.zeros(2, 64))                 # fx.py:80
    with pytest.raises(AssertionError):
        fx.apply_tremolo(x, torch.zeros(2, 64), mix=1.5)           # fx.py:21


def test_torch_ops_registered_cuda_only():
    """torch.ops.modfx.* exist and have no CPU kernel: the dispatcher itself refuses CPU tensors."""
    import torch
    import mod_extraction_b200._torch_ops  # noqa: F401
    assert hasattr(torch.ops.modfx, "flanger_chorus") and hasattr(torch.ops.modfx, "phaser")
    assert hasattr(torch.ops.modfx, "cnn_conv_pool_prelu") and hasattr(torch.ops.modfx, "cnn_head")
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.modfx.cnn_layernorm(torch.zeros(1, 2, 4, 4), True, 1e-5, False)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.modfx.interp_linear(torch.zeros(2, 8), 16, True)
