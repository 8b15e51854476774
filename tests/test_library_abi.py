"""CPU-side checks of the C-ABI boundary: the shared library loads without a GPU and exports every symbol that
include/knode_cosserat.h declares; argument validation works without touching a device."""
import ctypes as C
import os
import re

import pytest

import _kc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "knode_cosserat.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    L = _kc.lib()
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/knode_cosserat.h but not exported"
    assert sorted(_kc.exported_symbols()) == syms
    assert L.kc_version() >= 100


def test_struct_layout_matches_header():
    # 2 int32 + 5 double + 5*9 + 5*3 + (3+4+3+3) + 12 doubles
    assert C.sizeof(_kc.kc_rod_params) == 8 + 8 * (5 + 45 + 15 + 13 + 12)
    assert C.sizeof(_kc.kc_mlp) == 16 + 4 * 8


def test_argument_validation_needs_no_gpu():
    L = _kc.lib()
    p = _kc.kc_rod_params()
    p.N = 1
    rc = L.kc_ode_fwd(0, C.byref(p), None, 4, None, None, None, None, None, None, None)
    assert rc == -1 and b"N < 2" in L.kc_last_error()
    p.N = 10
    rc = L.kc_ode_fwd(7, C.byref(p), None, 4, None, None, None, None, None, None, None)
    assert rc == -1 and b"dtype" in L.kc_last_error()
    rc = L.kc_rollout_fwd(0, C.byref(p), None, 4, 10, None, None, None, 0.0, 0, 33, None, None, None, None, 0, None)
    assert rc == -1 and b"rows" in L.kc_last_error()
    assert L.kc_rollout_workspace_bytes(0, C.byref(p), None, 4096, 100) >= 4096 * 100 * 250 * 4
    one = C.c_void_p(8)
    rc = L.kc_estimate_state(1, C.byref(p), 0.4, 0.005, 1, 2, one, one, one, None)
    assert rc == -1 and b"T >= 3" in L.kc_last_error()
    p.N = 250
    rc = L.kc_estimate_state(1, C.byref(p), 0.4, 0.005, 1, 5, one, one, one, None)
    assert rc == -1 and b"too large" in L.kc_last_error()


def test_no_cpu_fallback():
    import torch
    import _ops
    from oracle import rod_oracle as O
    P = _kc.rod_params(O.RodParams())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ops.ode_fwd(P, None, torch.zeros(2, 19), torch.zeros(2, 19), torch.zeros(2, 6), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ops.estimate_state(P, 0.4, 0.005, torch.zeros(1, 4, 7, 10), torch.zeros(1, 4, 4))
