"""Memory-bound kernels (LN+modulate, gated residual, QK-RMSNorm+RoPE, GroupNorm+SiLU) vs the CPU oracle, fwd + bwd.
Tolerances: bf16 outputs max|y-ref|/max|ref| <= 2e-2 (observed ~4e-3 = bf16 rounding); gradient cosine >= 0.999."""
import pytest
import torch

from conftest import load_golden
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
TOL = 2e-2
COS = 0.999


def _r(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).to(torch.bfloat16)


def _leaf(t, dev=None, dtype=None):
    t = t.detach().clone()
    if dtype is not None:
        t = t.to(dtype)
    if dev is not None:
        t = t.to(dev)
    return t.requires_grad_(True)


@pytest.mark.parametrize("B,L,C,affine,mod", [(2, 37, 3072, False, True), (1, 130, 5120, False, True),
                                              (3, 50, 320, True, False), (2, 9, 1920, True, True),
                                              (1, 64, 1280, True, False)])
def test_ln_modulate_fwd_bwd(B, L, C, affine, mod):
    import b200vt.functional as Fn
    x = _r((B, L, C), 1, 2.0, 0.3)
    gamma = _r((C,), 2, 0.2, 1.0).float() if affine else None
    beta = _r((C,), 3, 0.2).float() if affine else None
    scale = _r((B, C), 4, 0.3).float() if mod else None
    shift = _r((B, C), 5, 0.3).float() if mod else None
    dy = _r((B, L, C), 6)
    # oracle (fp32 CPU)
    xr = _leaf(x, dtype=torch.float32)
    pr = [None if t is None else _leaf(t) for t in (gamma, beta, scale, shift)]
    yr = R.ln_modulate(xr, pr[0], pr[1], pr[2], pr[3], 1e-6)
    yr.backward(dy.float())
    # CUDA
    xc = _leaf(x, "cuda")
    pc = [None if t is None else _leaf(t, "cuda") for t in (gamma, beta, scale, shift)]
    yc = Fn.ln_modulate(xc, shift=pc[3], scale=pc[2], weight=pc[0], bias=pc[1], eps=1e-6)
    yc.backward(dy.cuda())
    assert R.max_rel_err(yc.float().cpu(), yr) < TOL
    assert R.cosine(xc.grad.float().cpu(), xr.grad) > COS
    assert R.max_rel_err(xc.grad.float().cpu(), xr.grad) < TOL
    for a, b_ in zip(pc, pr):
        if a is not None:
            assert R.cosine(a.grad.cpu(), b_.grad) > COS
            assert R.max_rel_err(a.grad.cpu(), b_.grad) < TOL


@pytest.mark.parametrize("B,L,C,gated", [(2, 33, 3072, True), (1, 77, 5120, True), (2, 10, 320, False)])
def test_gate_residual_fwd_bwd(B, L, C, gated):
    import b200vt.functional as Fn
    x, br, dy = _r((B, L, C), 1), _r((B, L, C), 2), _r((B, L, C), 3)
    gate = _r((B, C), 4).float() if gated else None
    xr, brr = _leaf(x, dtype=torch.float32), _leaf(br, dtype=torch.float32)
    gr = None if gate is None else _leaf(gate)
    yr = R.gate_residual(xr, brr, gr)
    yr.backward(dy.float())
    xc, brc = _leaf(x, "cuda"), _leaf(br, "cuda")
    gc = None if gate is None else _leaf(gate, "cuda")
    yc = Fn.gate_residual(xc, brc, gc)
    yc.backward(dy.cuda())
    assert R.max_rel_err(yc.float().cpu(), yr) < TOL
    assert R.max_rel_err(xc.grad.float().cpu(), xr.grad) < TOL
    assert R.max_rel_err(brc.grad.float().cpu(), brr.grad) < TOL
    if gated:
        assert R.cosine(gc.grad.cpu(), gr.grad) > COS and R.max_rel_err(gc.grad.cpu(), gr.grad) < TOL


def _rope_ref(x, w, cos, sin, per_head, eps):
    """oracle: hunyuan RMSNorm + apply_rotary_emb (per head) or wan RMSNorm over H*D + rope, in fp32."""
    B, L, H, D = x.shape
    if w is None:
        n = x
    elif per_head:
        n = R.hunyuan_rmsnorm(x, w, eps)
    else:
        n = R.wan_rmsnorm(x.reshape(B, L, H * D), w, eps).view(B, L, H, D)
    if cos is None:
        return n
    Lr = cos.shape[0]
    return torch.cat([R.hunyuan_apply_rotary_emb(n[:, :Lr], cos, sin), n[:, Lr:]], dim=1)


@pytest.mark.parametrize("B,L,H,D,per_head,use_w,Lr", [(1, 150, 3, 128, True, True, 120), (2, 70, 24, 128, True, True, 70),
                                                       (1, 61, 5, 128, False, True, 60), (2, 33, 2, 64, True, False, 33),
                                                       (1, 40, 40, 128, False, True, 40), (1, 20, 2, 128, True, True, 0),
                                                       # widths that are multiples of 1024 take the bulk-copy ring kernels:
                                                       (1, 90, 16, 64, True, True, 50),     # head dim 64, RoPE on a prefix
                                                       (2, 45, 8, 128, True, False, 45),    # no norm weight -> rotation only
                                                       (1, 700, 24, 128, True, True, 600),  # more tokens than ring stages
                                                       (1, 333, 16, 128, False, True, 300)])  # full-row norm, 2048 wide
def test_qk_rmsnorm_rope_fwd_bwd(B, L, H, D, per_head, use_w, Lr):
    import b200vt.functional as Fn
    x = _r((B, L, H, D), 1, 1.5)
    w = (_r((D if per_head else H * D,), 2, 0.1, 1.0).float()) if use_w else None
    cos = sin = None
    if Lr > 0:
        ang = torch.rand(Lr, D // 2, generator=torch.Generator().manual_seed(3)) * 6.28
        cos, sin = ang.cos().repeat_interleave(2, 1), ang.sin().repeat_interleave(2, 1)
    dy = _r((B, L, H, D), 4)
    xr = _leaf(x, dtype=torch.float32)
    wr = None if w is None else _leaf(w)
    yr = _rope_ref(xr, wr, cos, sin, per_head, 1e-6)
    yr.backward(dy.float())
    xc = _leaf(x, "cuda")
    wc = None if w is None else _leaf(w, "cuda")
    yc = Fn.qk_rmsnorm_rope(xc, wc, None if cos is None else cos.cuda(), None if sin is None else sin.cuda(),
                            per_head=per_head, eps=1e-6)
    yc.backward(dy.cuda())
    assert R.max_rel_err(yc.float().cpu(), yr) < TOL
    assert R.cosine(xc.grad.float().cpu(), xr.grad) > COS and R.max_rel_err(xc.grad.float().cpu(), xr.grad) < TOL
    if w is not None:
        assert R.cosine(wc.grad.cpu(), wr.grad) > COS and R.max_rel_err(wc.grad.cpu(), wr.grad) < TOL


def test_qk_rmsnorm_rope_strided_qkv_view_and_golden():
    import b200vt.functional as Fn
    g = load_golden("hunyuan_norm_rope")
    out = Fn.qk_rmsnorm_rope(g["xq"].cuda(), g["w"].cuda(), g["cos"].cuda(), g["sin"].cuda(), per_head=True, eps=g["eps"])
    assert R.max_rel_err(out.float().cpu(), g["rope_q"]) < TOL
    qkv = _r((1, 90, 3, 4, 128), 7).cuda()
    q = qkv.unbind(2)[1]
    w = torch.ones(128, device="cuda")
    ref = R.hunyuan_rmsnorm(q.float().cpu(), torch.ones(128), 1e-6)
    assert R.max_rel_err(Fn.qk_rmsnorm_rope(q, w, None, None).float().cpu(), ref) < TOL


def test_wan_rope_golden():
    import b200vt.functional as Fn
    g = load_golden("wan_ops")
    freqs = R.wan_freqs_table(128)
    cos, sin = R.wan_rope_cos_sin(g["grid"][0].tolist(), freqs)
    out = Fn.qk_rmsnorm_rope(g["q"][:1].cuda(), None, cos.cuda(), sin.cuda())
    assert R.max_rel_err(out.float().cpu(), g["roped"][:1]) < TOL
    y = Fn.qk_rmsnorm_rope(g["xr"].view(2, 60, 2, 128).cuda(), g["rms_w"].cuda(), None, None, per_head=False)
    assert R.max_rel_err(y.float().cpu().view(2, 60, 256), g["rms_out"]) < TOL
    y = Fn.ln_modulate(g["xr"].cuda(), eps=1e-6)
    assert R.max_rel_err(y.float().cpu(), g["ln_out"]) < TOL


@pytest.mark.parametrize("N,C,sp,dt,silu", [(4, 320, (40, 64), torch.bfloat16, True), (2, 640, (20, 32), torch.float32, True),
                                             (3, 1280, (5, 8), torch.bfloat16, False), (2, 128, (3, 7), torch.float32, True),
                                             (1, 320, (16, 9, 5), torch.bfloat16, True),
                                             # slabs split over a thread-block cluster (DSMEM reduction of the statistics):
                                             (2, 64, (16, 40, 64), torch.bfloat16, True),    # 164 KB slab -> 16 CTAs
                                             (2, 64, (8, 20, 32), torch.float32, True),      # fp32, 8 CTAs
                                             (1, 96, (3, 25, 56), torch.bfloat16, False),    # ragged last chunk
                                             (2, 960, (40, 64), torch.bfloat16, True)])      # 153 KB slab (skip concat)
def test_groupnorm_silu_fwd_bwd(N, C, sp, dt, silu):
    import b200vt.functional as Fn
    x = (_r((N, C) + sp, 1, 2.0, 0.7)).to(dt)
    gamma, beta = _r((C,), 2, 0.3, 1.0).float(), _r((C,), 3, 0.3).float()
    dy = _r((N, C) + sp, 4).to(dt)
    xr, gr, br = _leaf(x, dtype=torch.float32), _leaf(gamma), _leaf(beta)
    yr = R.groupnorm_silu(xr, gr, br, 32, 1e-5, silu)
    yr.backward(dy.float())
    xc, gc, bc = _leaf(x, "cuda"), _leaf(gamma, "cuda"), _leaf(beta, "cuda")
    yc = Fn.groupnorm_silu(xc, gc, bc, 32, 1e-5, silu)
    yc.backward(dy.cuda())
    tol = TOL if dt == torch.bfloat16 else 1e-4
    assert R.max_rel_err(yc.float().cpu(), yr) < tol
    assert R.cosine(xc.grad.float().cpu(), xr.grad) > COS and R.max_rel_err(xc.grad.float().cpu(), xr.grad) < tol * 2
    assert R.max_rel_err(gc.grad.cpu(), gr.grad) < tol * 2 and R.max_rel_err(bc.grad.cpu(), br.grad) < tol * 2


def test_groupnorm_golden():
    import b200vt.functional as Fn
    g = load_golden("lvdm_groupnorm")
    y = Fn.groupnorm_silu(g["x"].float().cuda(), g["weight"].cuda(), g["bias"].cuda(), 32, g["eps"], True)
    assert R.max_rel_err(y.cpu(), g["out_silu"]) < 1e-4
    y = Fn.groupnorm_silu(g["x"].cuda(), g["weight"].cuda(), g["bias"].cuda(), 32, g["eps"], False)
    assert R.max_rel_err(y.float().cpu(), g["out"]) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# full BASELINE sizes, where the CPU oracle is too slow: the same arithmetic written with torch fp32 ops on the GPU
# ---------------------------------------------------------------------------------------------------------------------
def test_ln_modulate_full_k1_size_matches_torch_fp32():
    """HunyuanVideo K1 activation (119 056 rows x 3072: the two-rows-per-iteration grid stride at full size): forward and input gradient against torch.layer_norm in fp32 on the GPU."""
    import torch.nn.functional as F

    import b200vt.functional as Fn
    g = torch.Generator(device="cuda").manual_seed(5)
    B, L, C = 1, 119056, 3072
    x = (torch.randn(B, L, C, device="cuda", generator=g) * 1.7 + 0.2).to(torch.bfloat16).requires_grad_(True)
    sc = torch.randn(B, C, device="cuda", generator=g) * 0.3
    sh = torch.randn(B, C, device="cuda", generator=g) * 0.3
    dy = torch.randn(B, L, C, device="cuda", generator=g).to(torch.bfloat16)
    y = Fn.ln_modulate(x, sh, sc, eps=1e-6)
    (gx,) = torch.autograd.grad(y, x, dy)
    xr = x.detach().float().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), eps=1e-6) * (1 + sc[:, None]) + sh[:, None]
    (gr,) = torch.autograd.grad(yr, xr, dy.float())
    assert float((y.float() - yr).abs().max() / yr.abs().max()) < TOL
    assert float(torch.nn.functional.cosine_similarity(gx.float().flatten(), gr.flatten(), dim=0)) > COS
    # every row was written (a skipped row would keep the allocator's stale bytes): per-row error, not only the global max
    row_err = (y.float() - yr).abs().amax(dim=-1) / yr.abs().amax(dim=-1)
    assert float(row_err.max()) < 4 * TOL


@pytest.mark.parametrize("shape", [(32, 320, 40, 64), (32, 960, 40, 64), (2, 320, 16, 40, 64), (2, 1280, 16, 10, 16)])
def test_groupnorm_silu_full_vc2_sizes_match_torch_fp32(shape):
    """VideoCrafter2 ResBlock / SpatialTransformer (4-D) and TemporalTransformer (5-D) GroupNorm inputs at batch 2: the
    bulk-copy kernels with 1-, 4- and 16-CTA clusters, forward + backward, against torch.group_norm + SiLU in fp32."""
    import torch.nn.functional as F

    import b200vt.functional as Fn
    g = torch.Generator(device="cuda").manual_seed(6)
    C = shape[1]
    x = (torch.randn(*shape, device="cuda", generator=g) * 2.0 + 0.5).to(torch.bfloat16).requires_grad_(True)
    w = (1 + 0.2 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    b = (0.2 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    dy = torch.randn(*shape, device="cuda", generator=g).to(torch.bfloat16)
    y = Fn.groupnorm_silu(x, w, b, 32, 1e-5, silu=True)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), dy)
    xr = x.detach().float().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.silu(F.group_norm(xr, 32, wr, br, 1e-5))
    gxr, gwr, gbr = torch.autograd.grad(yr, (xr, wr, br), dy.float())
    assert float((y.float() - yr).abs().max() / yr.abs().max()) < TOL
    assert float(F.cosine_similarity(gx.float().flatten(), gxr.flatten(), dim=0)) > COS
    assert float((gw - gwr).abs().max() / gwr.abs().max()) < TOL and float((gb - gbr).abs().max() / gbr.abs().max()) < TOL


def test_qk_rmsnorm_rope_full_k1_size_matches_torch_fp32():
    """HunyuanVideo K1: the q view of the fused QKV projection (118 800 tokens x 24 heads x 128, row stride 3 x 3072), per-head
    RMSNorm + RoPE, forward and backward through the bulk-copy ring kernels with their full-size grids (every CTA walks
    several tokens through the ring), against the same arithmetic in torch fp32 on the GPU."""
    import b200vt.functional as Fn
    g = torch.Generator(device="cuda").manual_seed(8)
    L, H, D = 118800, 24, 128
    qkv = torch.randn(1, L, 3, H, D, device="cuda", generator=g).to(torch.bfloat16)
    w = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    ang = torch.rand(L, D // 2, device="cuda", generator=g) * 6.28
    cos, sin = ang.cos().repeat_interleave(2, dim=1).contiguous(), ang.sin().repeat_interleave(2, dim=1).contiguous()
    dy = torch.randn(1, L, H, D, device="cuda", generator=g).to(torch.bfloat16)
    q = qkv[:, :, 0].detach().requires_grad_(True)
    wc = w.clone().requires_grad_(True)
    y = Fn.qk_rmsnorm_rope(q, wc, cos, sin, per_head=True, eps=1e-6)
    gq, gw = torch.autograd.grad(y, (q, wc), dy)
    qr = qkv[:, :, 0].detach().float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    n = qr * torch.rsqrt(qr.pow(2).mean(-1, keepdim=True) + 1e-6) * wr
    a, b = n.reshape(1, L, H, D // 2, 2).unbind(-1)
    rot = torch.stack([-b, a], dim=-1).flatten(3)
    yr = n * cos.view(1, L, 1, D) + rot * sin.view(1, L, 1, D)
    gqr, gwr = torch.autograd.grad(yr, (qr, wr), dy.float())
    assert float((y.float() - yr).abs().max() / yr.abs().max()) < TOL
    row_err = (y.float() - yr).abs().amax(dim=(2, 3)) / yr.abs().amax(dim=(2, 3))  # every token was written
    assert float(row_err.max()) < 4 * TOL
    assert float(torch.nn.functional.cosine_similarity(gq.float().flatten(), gqr.flatten(), dim=0)) > COS
    assert float((gq.float() - gqr).abs().max() / gqr.abs().max()) < 2 * TOL
    assert float((gw - gwr).abs().max() / gwr.abs().max()) < TOL


def test_registered_op_route_matches_eager_fast_path():
    """Under a multi-rank launch (and under tracing / dispatch modes) the ops run through torch.library's registered
    custom ops instead of the eager autograd.Function path (ops._make_eager): same results, and the registered route's
    stricter rules hold (no output may alias another output or an input) — forward and backward."""
    import b200vt.ops as ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(9)
    B, L, C = 2, 70, 256
    x = torch.randn(B, L, C, device=dev, generator=g).bfloat16()
    br = torch.randn(B, L, C, device=dev, generator=g).bfloat16()
    vecs = [torch.randn(B, C, device=dev, generator=g) * 0.1 for _ in range(6)]
    dy = torch.randn(B, L, C, device=dev, generator=g).bfloat16()

    def run(reg, split):
        ln = ops.ln_modulate_fwd.op if reg else ops.ln_modulate_fwd
        gr = ops.gate_residual_fwd.op if reg else ops.gate_residual_fwd
        leaves = [t.clone().requires_grad_(True) for t in (x, br, *vecs)]
        xx, bb, sc, sh, sc2, sh2, gt, gt2 = leaves
        if split:
            y = ln(xx, None, None, sc, sh, 1e-6, sc2, sh2, split)[0]
            z = gr(y, bb, gt, gt2, split)
        else:
            y = ln(xx, None, None, sc, sh, 1e-6)[0]
            z = gr(y, bb, gt)
        grads = torch.autograd.grad(z, [t for t in leaves if split or t is not sc2 and t is not sh2 and t is not gt2], dy)
        return [z.detach(), *grads]

    for split in (0, 25):
        a, b = run(True, split), run(False, split)
        for u, v in zip(a, b):
            assert float((u.float() - v.float()).abs().max()) <= 1e-2 * float(v.float().abs().max()) + 1e-6
    # joint q/k/v op through the registered route
    H, D, T = 2, 64, 10
    iq = torch.randn(1, L, 3, H, D, device=dev, generator=g).bfloat16().requires_grad_(True)
    tq = torch.randn(1, T, 3, H, D, device=dev, generator=g).bfloat16().requires_grad_(True)
    w = [(1 + 0.1 * torch.randn(D, device=dev, generator=g)).requires_grad_(True) for _ in range(4)]
    cs, sn = torch.randn(L, D, device=dev, generator=g).cos(), torch.randn(L, D, device=dev, generator=g).sin()
    outs = {}
    for reg in (True, False):
        f = ops.joint_qkv_fwd.op if reg else ops.joint_qkv_fwd
        q, k, v, _, _ = f(iq, tq, *w, cs, sn, 1e-6)
        gs = torch.autograd.grad([q, k, v], [iq, tq, *w], [torch.ones_like(q), torch.ones_like(k), torch.ones_like(v)])
        outs[reg] = [q.detach(), k.detach(), v.detach(), *gs]
    for u, v in zip(outs[True], outs[False]):
        assert float((u.float() - v.float()).abs().max()) <= 1e-3 * float(v.float().abs().max()) + 1e-6


@pytest.mark.parametrize("shape", [(4, 320, 12, 16), (3, 960, 5, 8), (2, 2560, 4, 4), (2, 64, 3, 5, 7), (2, 320, 16, 6, 8)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("silu", [True, False])
def test_groupnorm_channels_last_matches_fp32_reference_and_nchw_kernel(shape, dtype, silu):
    """Channels-last GroupNorm(+SiLU) kernels (csrc/groupnorm_nhwc.cu; torch.channels_last / channels_last_3d inputs)
    against torch's group_norm in fp32 and against the NCHW kernels, forward and backward; the output keeps the layout."""
    import torch.nn.functional as F
    import b200vt.functional as Fn
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(31)
    C = shape[1]
    fmt = torch.channels_last if len(shape) == 4 else torch.channels_last_3d
    x = (torch.randn(*shape, device=dev, generator=g) * 1.5 + 0.3).to(dtype)
    gw, gb = 1 + 0.1 * torch.randn(C, device=dev, generator=g), 0.1 * torch.randn(C, device=dev, generator=g)
    dy = torch.randn(*shape, device=dev, generator=g).to(dtype)
    x_cl = x.contiguous(memory_format=fmt).requires_grad_(True)
    wl, bl = gw.clone().requires_grad_(True), gb.clone().requires_grad_(True)
    y = Fn.groupnorm_silu(x_cl, wl, bl, 32, 1e-5, silu=silu)
    native = C <= (4096 if dtype == torch.bfloat16 else 2048)  # wider rows take the NCHW kernels (one layout copy)
    assert y.shape == x.shape and (y.is_contiguous(memory_format=fmt) or not native)
    y.backward(dy.contiguous(memory_format=fmt))
    assert x_cl.grad.is_contiguous(memory_format=fmt) or not native
    xr, wr, br = x.detach().float().clone().requires_grad_(True), gw.clone().requires_grad_(True), gb.clone().requires_grad_(True)
    ref = F.group_norm(xr, 32, wr, br, 1e-5)
    ref = F.silu(ref) if silu else ref
    ref.backward(dy.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-3
    rel = lambda a, b: float((a.float() - b.float()).abs().max() / b.float().abs().max())  # noqa: E731
    assert rel(y, ref) <= tol and rel(x_cl.grad, xr.grad) <= tol
    assert rel(wl.grad, wr.grad) <= tol and rel(bl.grad, br.grad) <= tol
    xn = x.detach().clone().requires_grad_(True)  # NCHW kernels on the same values
    yn = Fn.groupnorm_silu(xn, gw, gb, 32, 1e-5, silu=silu)
    yn.backward(dy)
    assert rel(y, yn) <= tol and rel(x_cl.grad, xn.grad) <= tol


@pytest.mark.parametrize("shape", [(3, 70, 2 * 1280), (2, 5, 7, 2 * 64), (1, 81920, 2 * 1280)])
def test_geglu_matches_exact_gelu_reference(shape):
    """Fused gated GELU (csrc/geglu.cu; lvdm GEGLU.forward, attention.py:527-529) against x * F.gelu(gate) in fp32, forward and
    backward; the last shape is the VideoCrafter2 level-0 feed-forward (batch 2 x 16 frames x 40 x 64 tokens, 8 * 320 wide)."""
    import torch.nn.functional as F
    import b200vt.ops as ops
    g = torch.Generator(device="cuda").manual_seed(17)
    xin = (torch.randn(*shape, device="cuda", generator=g) * 1.5).bfloat16().requires_grad_(True)
    dy = torch.randn(*shape[:-1], shape[-1] // 2, device="cuda", generator=g).bfloat16()
    y = ops.geglu_fwd(xin)
    y.backward(dy)
    xr = xin.detach().float().requires_grad_(True)
    a, gate = xr.chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    ref.backward(dy.float())
    rel = lambda u, v: float((u.float() - v).abs().max() / v.abs().max())  # noqa: E731
    assert rel(y, ref) <= 1e-2 and rel(xin.grad, xr.grad) <= 1e-2


@pytest.mark.parametrize("shape", [(4, 320, 12, 16), (3, 960, 5, 8), (2, 320, 16, 6, 8), (33, 64, 9, 7)])
@pytest.mark.parametrize("per_sample", [True, False])
@pytest.mark.parametrize("silu", [True, False])
def test_groupnorm_channels_last_with_folded_addend(shape, per_sample, silu):
    """GroupNorm(x + e[..., None, None]) with the per-channel term folded into the channels-last kernels (a convolution
    bias, (C,), and / or ResBlock's timestep embedding, (N, C)) against torch's group_norm of the explicit sum in fp32:
    output and input gradient; frozen affine parameters take the no-weight-gradient route of the backward."""
    import torch.nn.functional as F
    import b200vt.functional as Fn
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(77)
    N, C = shape[0], shape[1]
    fmt = torch.channels_last if len(shape) == 4 else torch.channels_last_3d
    x = (torch.randn(*shape, device=dev, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    e = torch.randn((N, C) if per_sample else (C,), device=dev, generator=g) * 0.7
    gw, gb = 1 + 0.1 * torch.randn(C, device=dev, generator=g), 0.1 * torch.randn(C, device=dev, generator=g)
    dy = torch.randn(*shape, device=dev, generator=g).to(torch.bfloat16)
    x_cl = x.contiguous(memory_format=fmt).requires_grad_(True)
    for frozen in (True, False):
        x_cl.grad = None
        wl, bl = gw.clone().requires_grad_(not frozen), gb.clone().requires_grad_(not frozen)
        y = Fn.groupnorm_silu(x_cl, wl, bl, 32, 1e-5, silu=silu, addend=e)
        assert y.is_contiguous(memory_format=fmt)
        y.backward(dy.contiguous(memory_format=fmt))
        xr = x.detach().float().clone().requires_grad_(True)
        wr, br = gw.clone().requires_grad_(True), gb.clone().requires_grad_(True)
        eb = e.view(*e.shape, *([1] * (len(shape) - 2))) if per_sample else e.view(1, C, *([1] * (len(shape) - 2)))
        ref = F.group_norm(xr + eb, 32, wr, br, 1e-5)
        ref = F.silu(ref) if silu else ref
        ref.backward(dy.float())
        rel = lambda a, b: float((a.float() - b.float()).abs().max() / b.float().abs().max())  # noqa: E731
        assert rel(y, ref) <= 2e-2 and rel(x_cl.grad, xr.grad) <= 2e-2
        if frozen:
            assert wl.grad is None and bl.grad is None
        else:
            assert rel(wl.grad, wr.grad) <= 2e-2 and rel(bl.grad, br.grad) <= 2e-2
    with pytest.raises(Exception):  # an addend that needs a gradient is refused, not silently dropped
        Fn.groupnorm_silu(x_cl, gw, gb, 32, 1e-5, silu=silu, addend=e.clone().requires_grad_(True))
    with pytest.raises(RuntimeError):  # NCHW activations: no addend
        Fn.groupnorm_silu(x.contiguous(), gw, gb, 32, 1e-5, silu=silu, addend=e)


def test_registered_op_route_groupnorm_and_geglu():
    """The GroupNorm ops (defaulted `addend` / `need_wgrad` arguments, channels-last and NCHW, trainable and frozen affine
    parameters) and the gated GELU through torch.library's registered ops (what a multi-rank launch uses) against the eager
    fast path."""
    import b200vt.ops as ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(19)
    x = torch.randn(3, 64, 6, 10, device=dev, generator=g).bfloat16()
    dy = torch.randn(3, 64, 6, 10, device=dev, generator=g).bfloat16()
    e = torch.randn(3, 64, device=dev, generator=g)
    gw, gb = 1 + 0.1 * torch.randn(64, device=dev, generator=g), 0.1 * torch.randn(64, device=dev, generator=g)
    for cl, addend, frozen in ((False, None, False), (True, None, False), (True, e, False), (True, e, True)):
        outs = {}
        for reg in (True, False):
            f = ops.groupnorm_silu_fwd.op if reg else ops.groupnorm_silu_fwd
            xx = (x.contiguous(memory_format=torch.channels_last) if cl else x.clone()).requires_grad_(True)
            w, b = gw.clone().requires_grad_(not frozen), gb.clone().requires_grad_(not frozen)
            args = (xx, w, b, 32, 1e-5, True) + ((addend,) if (addend is not None or not reg) else ())
            y = f(*args)[0]
            gs = torch.autograd.grad(y, [xx] if frozen else [xx, w, b], dy.contiguous(memory_format=torch.channels_last) if cl else dy)
            outs[reg] = [y.detach(), *gs]
        for u, v in zip(outs[True], outs[False]):
            assert float((u.float() - v.float()).abs().max()) <= 1e-3 * float(v.float().abs().max()) + 1e-6
    xin = torch.randn(50, 2 * 64, device=dev, generator=g).bfloat16()
    dz = torch.randn(50, 64, device=dev, generator=g).bfloat16()
    outs = {}
    for reg in (True, False):
        f = ops.geglu_fwd.op if reg else ops.geglu_fwd
        xi = xin.clone().requires_grad_(True)
        z = f(xi)
        outs[reg] = [z.detach(), torch.autograd.grad(z, xi, dz)[0]]
    for u, v in zip(outs[True], outs[False]):
        assert torch.equal(u, v)
