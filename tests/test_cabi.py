"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200vt.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols(experiments: bool = False):
    """Entry points the header declares for the product build (or, experiments=True, only inside #ifdef VT_EXPERIMENTS)."""
    text = open(os.path.join(ROOT, "include", "b200vt.h")).read()
    exp = re.search(r"#ifdef VT_EXPERIMENTS(.*?)#endif /\* VT_EXPERIMENTS \*/", text, flags=re.S)
    text = exp.group(1) if experiments else text.replace(exp.group(0), "")
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


def test_product_library_exports_no_experiment_hooks():
    """Microbenchmark / probe hooks and the earlier kernel variants live behind -DVT_EXPERIMENTS (tools/build_variant.sh);
    the in-tree product library must not carry them."""
    import b200vt._lib as L
    if os.environ.get("B200VT_LIB"):
        return  # an A/B build was selected on purpose
    exp = declared_symbols(experiments=True)
    assert set(exp) == set(L._EXPERIMENT_SIGS) and len(exp) == 4
    assert not [s for s in exp if hasattr(L.lib(), s)] and not L.has_experiments()


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


def test_eager_fast_path_keeps_registered_ops_for_tracing():
    """ops.* are thin callables (eager fast path, ops._make_eager) over the registered torch.library ops: under a dispatch
    mode (here FakeTensorMode, as torch.compile / make_fx use) they must route through the registered op and its fake
    kernel; on CPU tensors either path must raise (there is no CPU implementation)."""
    import pytest
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    import b200vt.ops as ops
    assert callable(ops.attn_fwd) and hasattr(torch.ops.b200vt, "attn_fwd")
    assert getattr(ops.attn_fwd, "op", ops.attn_fwd) is not None
    assert ops._eager_ok((torch.randn(2), None, 3))
    with FakeTensorMode():
        x = torch.empty(2, 5, 64, dtype=torch.bfloat16, device="cuda")
        sc = torch.empty(2, 64, dtype=torch.float32, device="cuda")
        assert not ops._eager_ok((x,))
        y, mean, rstd = ops.ln_modulate_fwd(x, None, None, sc, sc, 1e-6)
        assert tuple(y.shape) == (2, 5, 64) and y.dtype == torch.bfloat16 and tuple(mean.shape) == (10,)
        q = torch.empty(1, 40, 2, 64, dtype=torch.bfloat16, device="cuda")
        o, lse = ops.attn_fwd(q, q, q, None, None, None, 40, 40, 0.125)
        assert tuple(o.shape) == (1, 40, 2, 64) and tuple(lse.shape) == (1, 2, 40) and lse.dtype == torch.float32
    with pytest.raises(RuntimeError):
        ops.groupnorm_silu_fwd(torch.randn(2, 32, 4), None, None, 32, 1e-5, True)
