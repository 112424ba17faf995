"""Attention time of one VideoCrafter2 LoRA-finetune step (BASELINE.json configs[1]: bf16, batch 2, 320x512x16 frames):
every CrossAttention call the 3D-UNet makes (SURVEY.md §8 census: 16 SpatialTransformer + 17 TemporalTransformer
instances, `(B, heads, Nq, Nk, 64) x calls`), forward + backward, through the b200vt ops and through the reference's own
einsum / softmax / einsum core (lvdm/modules/attention.py:126-149) under bf16 autocast on the same GPU.
    python tools/bench_vc2_census.py [--iters 10]
Prints one JSON line per census entry and a total line (sum of per-call medians x calls = attention time of one UNet
forward + backward, without activation-checkpoint recompute). Device time from CUDA-graph replays (tools/_timing.py), 3
warm-ups, median; inputs rotate over enough copies to exceed the 126 MB L2."""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vt.functional as Fn  # noqa: E402
import b200vt.ops as ops  # noqa: E402
from _timing import device_time_ms  # noqa: E402

BATCH, FRAMES, D = 2, 16, 64
LEVELS = [  # (tokens per frame, heads, spatial-transformer instances, temporal-transformer instances)
    (2560, 5, 5, 5), (640, 10, 5, 5), (160, 20, 5, 5), (40, 20, 1, 1)]
CENSUS = []
for n, h, n_st, n_tt in LEVELS:
    CENSUS.append(("spatial_self", BATCH * FRAMES, n, n, h, n_st))
    CENSUS.append(("spatial_cross_text77", BATCH * FRAMES, n, 77, h, n_st))
    CENSUS.append(("temporal_self", BATCH * n, FRAMES, FRAMES, h, 2 * n_tt))  # attn1 and attn2 are both self-attention
CENSUS.append(("temporal_self_init_attn", BATCH * 2560, FRAMES, FRAMES, 8, 2))  # openaimodel3d.py:418-432


def timeit(fn, iters):
    return device_time_ms(fn, iters)


def ref_core(q, k, v, h, scale):
    """lvdm CrossAttention core on projected (B, N, h*D) tensors: '(b n (h d)) -> (b h) n d', einsum, softmax, einsum, back."""
    b, n, _ = q.shape
    def split(t):
        return t.view(b, t.shape[1], h, D).permute(0, 2, 1, 3).reshape(b * h, t.shape[1], D)
    qh, kh, vh = split(q), split(k), split(v)
    sim = torch.einsum("b i d, b j d -> b i j", qh, kh) * scale
    sim = sim.softmax(dim=-1)
    out = torch.einsum("b i j, b j d -> b i d", sim, vh)
    return out.view(b, h, n, D).permute(0, 2, 1, 3).reshape(b, n, h * D)


def ours_core(q, k, v, h, scale):
    b, n, _ = q.shape
    q4, k4, v4 = q.view(b, n, h, D), k.view(b, k.shape[1], h, D), v.view(b, v.shape[1], h, D)
    if n <= 32 and k.shape[1] <= 32:
        return ops.temporal_attn_fwd(q4, k4, v4, None, scale).reshape(b, n, h * D)
    return Fn.attention_blhd(q4, k4, v4, softmax_scale=scale).reshape(b, n, h * D)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(20230211)
    scale = D ** -0.5
    tot = {"ours_ms": 0.0, "torch_ms": 0.0, "flops": 0.0}
    for kind, B, nq, nk, h, calls in CENSUS:
        per = (B * nq * h * D * 2 + B * nk * h * D * 2) * 2 * 2  # q, o, dO-sized + k, v sized, fwd+bwd (rough working set)
        nbuf = max(2, min(16, int(math.ceil(300e6 / per))))
        sets = []
        for _ in range(nbuf):
            q = torch.randn(B, nq, h * D, device=dev, dtype=torch.bfloat16, generator=g).requires_grad_(True)
            k = torch.randn(B, nk, h * D, device=dev, dtype=torch.bfloat16, generator=g).requires_grad_(True)
            v = torch.randn(B, nk, h * D, device=dev, dtype=torch.bfloat16, generator=g).requires_grad_(True)
            do = torch.randn(B, nq, h * D, device=dev, dtype=torch.bfloat16, generator=g)
            sets.append((q, k, v, do))
        it = [0]

        def step(core):
            it[0] = (it[0] + 1) % nbuf
            q, k, v, do = sets[it[0]]
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = core(q, k, v, h, scale)
            torch.autograd.grad(out, (q, k, v), do.to(out.dtype))

        o_ms = timeit(lambda: step(ours_core), args.iters)
        t_ms = timeit(lambda: step(ref_core), args.iters)
        fl = 14.0 * B * h * nq * nk * D
        rec = {"entry": kind, "B": B, "heads": h, "Nq": nq, "Nk": nk, "D": D, "calls_per_unet_pass": calls,
               "ours_fwd_bwd_ms": round(o_ms, 4), "torch_einsum_fwd_bwd_ms": round(t_ms, 4),
               "speedup": round(t_ms / o_ms, 2), "ours_tflops": round(fl / o_ms / 1e9, 1)}
        print(json.dumps(rec), flush=True)
        tot["ours_ms"] += o_ms * calls
        tot["torch_ms"] += t_ms * calls
        tot["flops"] += fl * calls
        del sets
        torch.cuda.empty_cache()
    print(json.dumps({"entry": "TOTAL attention of one VideoCrafter2 UNet forward+backward (batch 2, 16 frames, 320x512)",
                      "ours_ms": round(tot["ours_ms"], 3), "torch_einsum_ms": round(tot["torch_ms"], 3),
                      "speedup": round(tot["torch_ms"] / tot["ours_ms"], 2),
                      "attention_tflop": round(tot["flops"] / 1e12, 3),
                      "ours_tflops": round(tot["flops"] / tot["ours_ms"] / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
