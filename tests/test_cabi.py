"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200vt.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200vt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t)\s+(vt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "vt_attn_fwd" in syms and "vt_version" in syms and len(syms) >= 10


def test_library_exports_every_declared_symbol():
    import b200vt._lib as L
    lib = L.lib()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in b200vt.h but not exported: {missing}"


def test_binding_table_matches_header():
    import b200vt._lib as L
    declared = set(declared_symbols())
    bound = set(L._SIGS)
    assert declared <= bound, f"no ctypes signature for {sorted(declared - bound)}"


def test_version_and_error_string_without_gpu():
    import b200vt._lib as L
    lib = L.lib()
    assert lib.vt_version() >= 100
    # argument validation happens before any CUDA call: NULL pointers are rejected with VT_ERR_NULL (-5)
    rc = lib.vt_attn_fwd(None, None, None, None, None, None, None, None, None, 1, 1, 1, 1, 128, None, None, 0, 0, 0,
                         None, ctypes.c_float(1.0), None)
    assert rc == -5
    assert "NULL" in L.last_error()


def test_ops_registered_without_gpu():
    import torch
    import b200vt.ops  # noqa: F401
    for name in ("attn_fwd", "attn_bwd", "ln_modulate_fwd", "gate_residual_fwd", "qk_rmsnorm_rope_fwd",
                 "groupnorm_silu_fwd"):
        assert hasattr(torch.ops.b200vt, name)


def test_no_cpu_fallback():
    import pytest
    import torch
    import b200vt.ops as ops
    q = torch.zeros(1, 128, 1, 128, dtype=torch.bfloat16)
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.attn_fwd(q, q, q, None, None, None, 128, 128, 0.1)
