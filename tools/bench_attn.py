"""Kernel micro-benchmark: attention fwd / fwd+bwd TFLOP/s at the BASELINE shapes, ours vs library kernels on the same
GPU (flash-attn 2.8.3 FA2, torch SDPA). CUDA-event timing, warm-up, median. Usage:
    python tools/bench_attn.py [--shapes k1,k2,k3,k4] [--iters 5] [--libs] [--bwd]"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = {  # name: (B, Lq, Lk, H, D)
    "k1": (1, 119056, 119056, 24, 128),   # HunyuanVideo 720x1280x129
    "k2": (1, 32760, 32760, 40, 128),     # Wan2.1-14B 480x832x81 self-attention
    "k2x": (1, 32760, 512, 40, 128),      # Wan cross-attention
    "k3": (1, 17776, 17776, 30, 64),      # CogVideoX-2B
    "k4": (32, 2560, 2560, 5, 64),        # VideoCrafter2 spatial self-attention, level 0, batch 2
    "k4x": (32, 2560, 77, 5, 64),         # VideoCrafter2 cross-attention
    "s8k": (1, 8192, 8192, 24, 128),
    "s16k": (1, 16384, 16384, 24, 128),
}


def timeit(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="s8k,k4,k3,k2,k1")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--libs", action="store_true")
    ap.add_argument("--bwd", action="store_true")
    args = ap.parse_args()
    import b200vt.ops as ops
    for name in args.shapes.split(","):
        B, Lq, Lk, H, D = SHAPES[name]
        g = torch.Generator(device="cuda").manual_seed(20230211)
        q = torch.randn(B, Lq, H, D, device="cuda", dtype=torch.bfloat16, generator=g)
        k = torch.randn(B, Lk, H, D, device="cuda", dtype=torch.bfloat16, generator=g)
        v = torch.randn(B, Lk, H, D, device="cuda", dtype=torch.bfloat16, generator=g)
        do = torch.randn(B, Lq, H, D, device="cuda", dtype=torch.bfloat16, generator=g)
        scale = 1 / math.sqrt(D)
        flops_f = 4.0 * B * H * Lq * Lk * D
        rec = {"shape": name, "B": B, "Lq": Lq, "Lk": Lk, "H": H, "D": D}
        o, lse = ops.attn_fwd(q, k, v, None, None, None, Lq, Lk, scale)
        ms = timeit(lambda: ops.attn_fwd(q, k, v, None, None, None, Lq, Lk, scale), args.iters)
        rec["ours_fwd_ms"], rec["ours_fwd_tflops"] = round(ms, 3), round(flops_f / ms / 1e9, 1)
        if args.bwd:
            ms = timeit(lambda: ops.attn_bwd(do, q, k, v, o, lse, None, None, None, Lq, Lk, scale), args.iters)
            rec["ours_bwd_ms"], rec["ours_bwd_tflops"] = round(ms, 3), round(2.5 * flops_f / ms / 1e9, 1)
            rec["ours_fwdbwd_tflops"] = round(3.5 * flops_f / (ms + rec["ours_fwd_ms"]) / 1e9, 1)
        if args.libs:
            try:
                from flash_attn import flash_attn_func
                ms = timeit(lambda: flash_attn_func(q, k, v), args.iters)
                rec["fa2_fwd_ms"], rec["fa2_fwd_tflops"] = round(ms, 3), round(flops_f / ms / 1e9, 1)
                if args.bwd:
                    qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
                    og = flash_attn_func(qg, kg, vg)
                    ms = timeit(lambda: torch.autograd.grad(og, (qg, kg, vg), do, retain_graph=True), args.iters)
                    rec["fa2_bwd_ms"], rec["fa2_bwd_tflops"] = round(ms, 3), round(2.5 * flops_f / ms / 1e9, 1)
            except Exception as e:  # noqa: BLE001
                rec["fa2_error"] = repr(e)[:120]
            try:
                qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
                ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt), args.iters)
                rec["sdpa_fwd_ms"], rec["sdpa_fwd_tflops"] = round(ms, 3), round(flops_f / ms / 1e9, 1)
                from torch.nn.attention import SDPBackend, sdpa_kernel
                with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
                    ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt), args.iters)
                rec["cudnn_fwd_ms"], rec["cudnn_fwd_tflops"] = round(ms, 3), round(flops_f / ms / 1e9, 1)
                if args.bwd:
                    qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (qt, kt, vt))
                    with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
                        og = torch.nn.functional.scaled_dot_product_attention(qg, kg, vg)
                        dot = do.transpose(1, 2)
                        ms = timeit(lambda: torch.autograd.grad(og, (qg, kg, vg), dot, retain_graph=True), args.iters)
                    rec["cudnn_bwd_ms"], rec["cudnn_bwd_tflops"] = round(ms, 3), round(2.5 * flops_f / ms / 1e9, 1)
            except Exception as e:  # noqa: BLE001
                rec["sdpa_error"] = repr(e)[:120]
        print(json.dumps(rec), flush=True)
        del q, k, v, do, o, lse
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
