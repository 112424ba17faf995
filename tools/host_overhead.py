"""Host-side cost of one op call (dispatcher + ctypes + output allocation), measured on inputs small enough that the GPU is
never the bottleneck: N back-to-back calls, one synchronize, wall time / N. Ours next to the torch ops the reference
launches for the same arithmetic.    python tools/host_overhead.py [--n 3000]"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vt.functional as Fn  # noqa: E402
import b200vt.ops as ops  # noqa: E402


def per_call_us(fn, n):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3000)
    a = ap.parse_args()
    dev = "cuda"
    x = torch.randn(2, 64, 8, 8, device=dev, dtype=torch.bfloat16)
    w, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    xl = torch.randn(1, 64, 1024, device=dev, dtype=torch.bfloat16)
    sc = torch.zeros(1, 1024, device=dev)
    q = torch.randn(1, 128, 2, 64, device=dev, dtype=torch.bfloat16)
    qt = torch.randn(8, 16, 2, 64, device=dev, dtype=torch.bfloat16)
    rows = {
        "groupnorm_silu": (lambda: Fn.groupnorm_silu(x, w, b, 32, 1e-5, silu=True),
                           lambda: F.silu(F.group_norm(x.float(), 32, w, b, 1e-5).to(x.dtype))),
        "ln_modulate": (lambda: Fn.ln_modulate(xl, sc, sc, eps=1e-6),
                        lambda: (F.layer_norm(xl.float(), (1024,), eps=1e-6) * (1 + sc[:, None]) + sc[:, None]).to(xl.dtype)),
        "gate_residual": (lambda: Fn.gate_residual(xl, xl, sc), lambda: xl + xl * sc[:, None].to(xl.dtype)),
        "attention_blhd": (lambda: Fn.attention_blhd(q, q, q),
                           lambda: F.scaled_dot_product_attention(q.transpose(1, 2), q.transpose(1, 2), q.transpose(1, 2))),
        "temporal_attn": (lambda: ops.temporal_attn_fwd(qt, qt, qt, None, 0.125),
                          lambda: torch.einsum("bhij,bhjd->bhid", torch.einsum("bihd,bjhd->bhij", qt, qt).softmax(-1),
                                               qt.transpose(1, 2))),
        "torch.empty_like (floor)": (lambda: torch.empty_like(xl), lambda: torch.empty_like(xl)),
    }
    for name, (ours, ref) in rows.items():
        print(json.dumps({"op": name, "ours_host_us_per_call": round(per_call_us(ours, a.n), 1),
                          "torch_ops_host_us_per_call": round(per_call_us(ref, a.n), 1)}), flush=True)


if __name__ == "__main__":
    main()
