"""ctypes binding of libb200vt.so (the C ABI in include/b200vt.h). Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200VT_LIB selects another build of the same library (A/B kernel experiments); the default is the in-tree build.
LIB_PATH = os.environ.get("B200VT_LIB") or os.path.join(_HERE, "libb200vt.so")

_lib = None
_lock = threading.Lock()

c_i64p = C.POINTER(C.c_int64)
vp = C.c_void_p

# name -> argtypes (restype is always int). Kept textually close to include/b200vt.h.
_SIGS = {
    "vt_version": [],
    "vt_last_error": [C.c_char_p, C.c_size_t],
    "vt_init": [C.c_int],
    "vt_debug_watchdog": [C.POINTER(C.c_uint32)],
    "vt_memcpy2d_async": [vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp],
    "vt_debug_set_trace": [vp],
    "vt_profile_enable": [C.c_int],
    "vt_profile_read": [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)],
    "vt_attn_fwd": [vp, vp, vp, vp, vp, c_i64p, c_i64p, c_i64p, c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                    vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_float, vp],
    "vt_attn_fwd_scatter": [vp, vp, vp, vp, vp, c_i64p, c_i64p, c_i64p, c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, vp,
                            C.c_float, C.POINTER(vp), C.c_int, C.c_int, c_i64p, vp],
    "vt_attn_bwd_workspace_bytes": [C.c_int, C.c_int, C.c_int, C.c_int],
    "vt_attn_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp] + [c_i64p] * 8 + [C.c_int] * 5 + [vp, vp, C.c_int, C.c_int,
                    C.c_int, vp, C.c_float, vp, C.c_int64, vp],
    "vt_temporal_attn_fwd": [vp, vp, vp, vp, vp, c_i64p, c_i64p, c_i64p, c_i64p, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_float, vp],
    "vt_temporal_attn_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, c_i64p, c_i64p, c_i64p, c_i64p, C.c_int, C.c_int, C.c_int,
                             C.c_int, C.c_float, vp],
    "vt_qk_rmsnorm_rope_fwd": [vp, vp, vp, vp, vp, vp, c_i64p, c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_float, vp],
    "vt_qk_rmsnorm_rope_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, c_i64p, c_i64p, c_i64p, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_int, vp],
    "vt_ln_modulate_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, vp],
    "vt_ln_modulate_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp],
    "vt_gate_residual_fwd": [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp],
    "vt_gate_residual_bwd": [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp],
    "vt_groupnorm_silu_fwd": [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                              vp],
    "vt_groupnorm_silu_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              vp],
    "vt_geglu_fwd": [vp, vp, C.c_int64, C.c_int, vp],
    "vt_geglu_bwd": [vp, vp, vp, C.c_int64, C.c_int, vp],
    "vt_groupnorm_nhwc_workspace_bytes": [C.c_int, C.c_int],
    "vt_groupnorm_silu_nhwc_fwd": [vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_int, C.c_int, vp],
    "vt_groupnorm_silu_nhwc_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, vp],
}
# Present only in -DVT_EXPERIMENTS builds of the library (tools/build_variant.sh + B200VT_LIB); see include/b200vt.h.
_EXPERIMENT_SIGS = {
    "vt_umma_rate": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp],
    "vt_tma_reduce_rate": [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp],
    "vt_tma_mixed_rate": [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp],
    "vt_umma_probe": [vp, vp, vp, C.c_int, C.c_int, C.c_int] + [C.c_uint32] * 6 + [vp],
}
_RESTYPE = {"vt_attn_bwd_workspace_bytes": C.c_int64, "vt_groupnorm_nhwc_workspace_bytes": C.c_int64}


class B200VTError(RuntimeError):
    """A libb200vt entry point returned a negative VT_ERR_* code."""

    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.fn, self.code, self.msg = fn, code, msg


def lib() -> C.CDLL:
    """Load the shared library once. No fallback: a missing library is an error, not a slow path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc with sm_100a support). b200vt has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in {**_SIGS, **_EXPERIMENT_SIGS}.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                continue  # entry point not built yet; calling it raises in call()
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = handle
        return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().vt_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def watchdog() -> tuple[int, int, int, int]:
    arr = (C.c_uint32 * 4)()
    lib().vt_debug_watchdog(arr)
    return tuple(int(v) for v in arr)


K_ATTN_FWD, K_ATTN_BWD, K_ATTN_BWD_DELTA, K_ATTN_BWD_DQ = 0, 1, 2, 3


def profile_enable(on: bool) -> None:
    call("vt_profile_enable", int(on))


def profile_read(kernel_id: int) -> tuple[float, int]:
    """(total device ms, launches) of one kernel id since profile_enable(True); waits for the recorded events."""
    ms, n = C.c_double(0), C.c_int64(0)
    call("vt_profile_read", kernel_id, C.byref(ms), C.byref(n))
    return ms.value, n.value


def call(name: str, *args):
    """Invoke an int-returning entry point and raise B200VTError on a negative code."""
    handle = lib()
    fn = getattr(handle, name, None)
    if fn is None:
        raise B200VTError(name, -6, "entry point not present in libb200vt.so")
    rc = fn(*args)
    if rc < 0:
        raise B200VTError(name, rc, last_error())
    return rc


def exported_symbols() -> list[str]:
    return [n for n in _SIGS if hasattr(lib(), n)]


def has_experiments() -> bool:
    """True when the loaded library was built with -DVT_EXPERIMENTS (microbenchmark hooks, earlier kernel variants)."""
    return hasattr(lib(), "vt_umma_probe")


def strides3(t, dims=(0, 1, 2)):
    """(b,l,h) element strides of a 4-D tensor as a ctypes int64[3]."""
    s = t.stride()
    return (C.c_int64 * 3)(s[dims[0]], s[dims[1]], s[dims[2]])
