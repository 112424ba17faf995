"""Achieved HBM bandwidth of the memory-bound kernels around attention at the BASELINE shapes, next to
the same arithmetic written with stock PyTorch ops on the same GPU (what the reference's block code launches).
Algorithmic bytes per element are the ones DESIGN.md states; the roofline is the measured copy bandwidth in
MEASURED_PEAKS.json. Device time per call from CUDA-graph replays (tools/_timing.py: no host launch path inside the timed
region), 3 warm-ups, median of `--iters`; tensors are far larger than the 126 MB L2 or rotate over enough copies to be.
    python tools/bench_rowwise.py [--iters 10]"""
import argparse
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vt.functional as Fn  # noqa: E402
import b200vt.ops as ops  # noqa: E402
from _timing import device_time_ms  # noqa: E402


def timeit(fn, iters):
    return device_time_ms(fn, iters)


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-torch", action="store_true", help="skip the stock-PyTorch comparison columns")
    args = ap.parse_args()
    run(args.iters, with_torch=not args.no_torch, emit=True)


def run(iters: int = 10, with_torch: bool = True, emit: bool = True):
    """Returns the list of records (also printed when emit). with_torch=False skips the stock-PyTorch comparison arm
    (bench.py's `rowwise` key: our kernels only)."""
    import types
    args = types.SimpleNamespace(iters=iters)
    records = []
    dev = "cuda"
    peak, src = peak_gbs()
    g = torch.Generator(device=dev).manual_seed(20230211)

    def rn(*shape, dtype=torch.bfloat16):
        return torch.randn(*shape, device=dev, dtype=torch.float32, generator=g).to(dtype)

    def report(name, shape, nbytes, ours_ms, torch_ms, note=""):
        rec = {"kernel": name, "shape": list(shape), "algorithmic_MB": round(nbytes / 1e6, 1), "ours_ms": round(ours_ms, 4),
               "ours_GBps": round(nbytes / ours_ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ours_ms / 1e6 / peak, 3),
               "torch_ops_ms": round(torch_ms, 4), "speedup_vs_torch_ops": round(torch_ms / ours_ms, 2),
               "peak_GBps": peak, "peak_kind": src, "timing": device_time_ms.last_mode}
        if note:
            rec["note"] = note
        if not with_torch:
            rec.pop("torch_ops_ms"), rec.pop("speedup_vs_torch_ops")
        records.append(rec)
        if emit:
            print(json.dumps(rec), flush=True)

    def timeit_ref(fn, iters):
        return timeit(fn, iters) if with_torch else float("nan")

    # ---- LayerNorm + modulate (Hunyuan K1 image stream; Wan C5 with the fp32 residual stream) -----------------------
    for tag, (B, L, C), xdt in (("hunyuan_k1", (1, 119056, 3072), torch.bfloat16), ("wan_c5_fp32_stream", (1, 32760, 5120), torch.float32)):
        x = rn(B, L, C, dtype=xdt).requires_grad_(True)
        sc, sh = rn(B, C, dtype=torch.float32) * 0.1, rn(B, C, dtype=torch.float32) * 0.1
        dy = rn(B, L, C)
        esz = x.element_size()
        y = Fn.ln_modulate(x, sh, sc, eps=1e-6)
        f_ms = timeit(lambda: Fn.ln_modulate(x, sh, sc, eps=1e-6), args.iters)
        b_ms = timeit(lambda: torch.autograd.grad(y, x, dy, retain_graph=True), args.iters)

        def ref():
            return (F.layer_norm(x.float(), (C,), eps=1e-6) * (1 + sc[:, None]) + sh[:, None]).to(torch.bfloat16)
        yr = ref() if with_torch else None
        rf_ms = timeit_ref(ref, args.iters)
        rb_ms = timeit_ref(lambda: torch.autograd.grad(yr, x, dy, retain_graph=True), args.iters)
        n = B * L * C
        report(f"ln_modulate_fwd[{tag}]", (B, L, C), n * (esz + 2), f_ms, rf_ms)
        report(f"ln_modulate_bwd[{tag}]", (B, L, C), n * (2 + esz + esz), b_ms, rb_ms)
        del x, y, yr, dy

    # ---- gated residual -------------------------------------------------------------------------------------------
    for tag, (B, L, C), xdt in (("hunyuan_k1", (1, 119056, 3072), torch.bfloat16), ("wan_c5_fp32_stream", (1, 32760, 5120), torch.float32)):
        x, br, gate = rn(B, L, C, dtype=xdt), rn(B, L, C), rn(B, C, dtype=torch.float32)
        f_ms = timeit(lambda: Fn.gate_residual(x, br, gate), args.iters)
        rf_ms = timeit_ref(lambda: x + br * gate[:, None].to(br.dtype if xdt == torch.bfloat16 else torch.float32), args.iters)
        n = B * L * C
        report(f"gate_residual_fwd[{tag}]", (B, L, C), n * (2 * x.element_size() + 2), f_ms, rf_ms)
        xg, bg, gg = x.detach().requires_grad_(True), br.detach().requires_grad_(True), gate.detach().requires_grad_(True)
        yg = Fn.gate_residual(xg, bg, gg)
        dyg = rn(B, L, C, dtype=xdt)
        b_ms = timeit(lambda: torch.autograd.grad(yg, (xg, bg, gg), dyg, retain_graph=True), args.iters)
        yr_ = (xg + bg * gg[:, None].to(bg.dtype if xdt == torch.bfloat16 else torch.float32)) if with_torch else None
        rb_ms = timeit_ref(lambda: torch.autograd.grad(yr_, (xg, bg, gg), dyg, retain_graph=True), args.iters)
        # dy read, branch read (for dgate), dbranch write; dx = dy is returned without a copy
        report(f"gate_residual_bwd[{tag}]", (B, L, C), n * (x.element_size() + 2 + 2), b_ms, rb_ms)
        del x, br, xg, bg, yg, dyg, yr_

    # ---- fused QK-RMSNorm + RoPE on the strided q view of a fused QKV projection (Hunyuan K1) -------------------------
    B, L, H, D = 1, 118800, 24, 128
    qkv = rn(B, L, 3, H, D)
    w = (1 + 0.1 * rn(D, dtype=torch.float32))
    ang = torch.rand(L, D // 2, device=dev, generator=g) * 6.28
    cos, sin = ang.cos().repeat_interleave(2, dim=1).contiguous(), ang.sin().repeat_interleave(2, dim=1).contiguous()
    q = qkv[:, :, 0]
    f_ms = timeit(lambda: Fn.qk_rmsnorm_rope(q, w, cos, sin), args.iters)

    def ref_rope():
        xf = q.float()
        n_ = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(q.dtype) * w.to(q.dtype)
        nf = n_.float()
        xr, xi = nf.reshape(*nf.shape[:-1], -1, 2).unbind(-1)
        rot = torch.stack([-xi, xr], dim=-1).flatten(3)
        return (nf * cos.view(1, L, 1, D) + rot * sin.view(1, L, 1, D)).to(q.dtype)
    rf_ms = timeit_ref(ref_rope, max(3, args.iters // 3))
    n = B * L * H * D
    report("qk_rmsnorm_rope_fwd[hunyuan_k1, strided q of fused qkv]", (B, L, H, D), n * 4 + 2 * L * D * 4, f_ms, rf_ms,
           note="reference = hunyuan RMSNorm (fp32 temporaries) + apply_rotary_emb (rotate_half stack/flatten) as torch ops")
    qg = qkv[:, :, 0].detach().requires_grad_(True)   # strided leaf: the backward reads x in place, like the block does
    yq = Fn.qk_rmsnorm_rope(qg, w, cos, sin)
    dyq = rn(B, L, H, D)
    b_ms = timeit(lambda: torch.autograd.grad(yq, qg, dyq, retain_graph=True), args.iters)
    qr = qkv[:, :, 0].detach().requires_grad_(True)

    def ref_rope_g():
        xf = qr.float()
        n_ = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(qr.dtype) * w.to(qr.dtype)
        nf = n_.float()
        xr, xi = nf.reshape(*nf.shape[:-1], -1, 2).unbind(-1)
        rot = torch.stack([-xi, xr], dim=-1).flatten(3)
        return (nf * cos.view(1, L, 1, D) + rot * sin.view(1, L, 1, D)).to(qr.dtype)
    yr_ = ref_rope_g() if with_torch else None
    rb_ms = timeit_ref(lambda: torch.autograd.grad(yr_, qr, dyq, retain_graph=True), max(3, args.iters // 3))
    report("qk_rmsnorm_rope_bwd[hunyuan_k1, strided q of fused qkv]", (B, L, H, D), n * 6 + 2 * L * D * 4 + B * L * H * 4,
           b_ms, rb_ms)
    del qkv, q, qg, yq, dyq, qr, yr_

    # ---- GroupNorm(32) + SiLU (VideoCrafter2 UNet, batch 2 x 16 frames): ResBlock / SpatialTransformer 4-D inputs and the
    # TemporalTransformer's 5-D input (N = 2 samples, slabs of (C/32) * t*h*w elements). These tensors are smaller than the
    # 126 MB L2, so every call takes the next of enough distinct copies to cycle through >= 512 MB.
    for shape in ((32, 320, 40, 64), (32, 960, 40, 64), (32, 1280, 10, 16), (2, 320, 16, 40, 64), (2, 1280, 16, 10, 16)):
        C = shape[1]
        nbuf = max(2, int(math.ceil(512e6 / (math.prod(shape) * 4))))
        xs = [rn(*shape).requires_grad_(True) for _ in range(nbuf)]
        dys = [rn(*shape) for _ in range(nbuf)]
        gw, gb = 1 + 0.1 * rn(C, dtype=torch.float32), 0.1 * rn(C, dtype=torch.float32)
        it = [0]

        def nxt():
            it[0] = (it[0] + 1) % nbuf
            return it[0]
        f_ms = timeit(lambda: Fn.groupnorm_silu(xs[nxt()], gw, gb, 32, 1e-5, silu=True), args.iters * 2)
        rf_ms = timeit_ref(lambda: F.silu(F.group_norm(xs[nxt()].float(), 32, gw, gb, 1e-5).to(torch.bfloat16)), args.iters * 2)
        ys = [Fn.groupnorm_silu(x, gw, gb, 32, 1e-5, silu=True) for x in xs]
        yrs = [F.silu(F.group_norm(x.float(), 32, gw, gb, 1e-5).to(torch.bfloat16)) for x in xs] if with_torch else None

        def bwd(outs):
            i = nxt()
            return torch.autograd.grad(outs[i], xs[i], dys[i], retain_graph=True)
        b_ms = timeit(lambda: bwd(ys), args.iters * 2)
        rb_ms = timeit_ref(lambda: bwd(yrs), args.iters * 2)
        n = math.prod(shape)
        note = (f"reference = GroupNormSpecific (x.float() -> group_norm -> type(x.dtype)) + SiLU; {nbuf} rotating buffers "
                "(tensor < L2)")
        report("groupnorm_silu_fwd[vc2]", shape, n * 4, f_ms, rf_ms, note=note)
        report("groupnorm_silu_bwd[vc2]", shape, n * 6, b_ms, rb_ms, note=note)
        # the same tensors channels-last (patch.lvdm_channels_last flow): two sweeps per direction, the second from L2
        fmt = torch.channels_last if len(shape) == 4 else torch.channels_last_3d
        xc = [x.detach().contiguous(memory_format=fmt).requires_grad_(True) for x in xs]
        dc = [d.contiguous(memory_format=fmt) for d in dys]
        f_ms = timeit(lambda: Fn.groupnorm_silu(xc[nxt()], gw, gb, 32, 1e-5, silu=True), args.iters * 2)
        yc = [Fn.groupnorm_silu(x, gw, gb, 32, 1e-5, silu=True) for x in xc]

        def bwd_c():
            i = nxt()
            return torch.autograd.grad(yc[i], xc[i], dc[i], retain_graph=True)
        b_ms = timeit(bwd_c, args.iters * 2)
        report("groupnorm_silu_fwd[vc2, channels_last]", shape, n * 4, f_ms, float("nan"), note="algorithmic bytes: one read + one write")
        report("groupnorm_silu_bwd[vc2, channels_last]", shape, n * 6, b_ms, float("nan"), note="algorithmic bytes: two reads + one write")
        del xs, dys, ys, yrs, xc, dc, yc

    # ---- temporal micro-attention, N = 16 frames (VideoCrafter2 level 0: 2 x 40 x 64 positions, 5 heads) --------------
    Bt, N, H, D = 5120, 16, 5, 64
    q, k, v = rn(Bt, N, H, D), rn(Bt, N, H, D), rn(Bt, N, H, D)
    sc = 1 / math.sqrt(D)
    f_ms = timeit(lambda: ops.temporal_attn_fwd(q, k, v, None, sc), args.iters)
    qt, kt, vt = (t.permute(0, 2, 1, 3).reshape(Bt * H, N, D) for t in (q, k, v))

    def ref_t():  # lvdm einsum / softmax / einsum (attention.py:128-144)
        s = torch.einsum("b i d, b j d -> b i j", qt, kt) * sc
        return torch.einsum("b i j, b j d -> b i d", s.softmax(dim=-1), vt)
    rf_ms = timeit_ref(ref_t, args.iters)
    report("temporal_attn_fwd[vc2 level 0]", (Bt, N, H, D), q.numel() * 2 * 4, f_ms, rf_ms)
    do = rn(Bt, N, H, D)
    b_ms = timeit(lambda: ops.temporal_attn_bwd(do, q, k, v, None, sc), args.iters)
    qr, kr, vr = (t.detach().clone().requires_grad_(True) for t in (qt, kt, vt))
    dot = do.permute(0, 2, 1, 3).reshape(Bt * H, N, D)
    s_ = torch.einsum("b i d, b j d -> b i j", qr, kr) * sc
    out_r = torch.einsum("b i j, b j d -> b i d", s_.softmax(dim=-1), vr)
    rb_ms = timeit_ref(lambda: torch.autograd.grad(out_r, (qr, kr, vr), dot, retain_graph=True), args.iters)
    report("temporal_attn_bwd[vc2 level 0]", (Bt, N, H, D), q.numel() * 2 * 7, b_ms, rb_ms)
    return records


if __name__ == "__main__":
    main()
