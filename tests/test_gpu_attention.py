"""CUDA attention forward/backward (through torch.library -> C ABI) against the CPU oracle and the golden fixtures.
Tolerances: bf16 output max|y-ref|/max|ref| <= 2e-2 vs the fp32 oracle; gradient cosine >= 0.999."""
import math

import pytest
import torch

from conftest import load_golden
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(torch.bfloat16)


def _run(q, k, v, **kw):
    import b200vt.functional as Fn
    return Fn.attention_blhd(q.cuda(), k.cuda(), v.cuda(), **kw)


@pytest.mark.parametrize("B,Lq,Lk,H,D", [
    (1, 128, 128, 1, 128), (1, 256, 256, 2, 128), (2, 200, 333, 2, 128), (1, 1000, 77, 3, 64), (2, 40, 40, 5, 64),
    (1, 640, 640, 2, 64), (1, 513, 1025, 1, 128), (3, 16, 16, 2, 64), (1, 2560, 2560, 2, 64),
])
def test_fwd_matches_oracle(B, Lq, Lk, H, D):
    q, k, v = _rand((B, Lq, H, D), 1), _rand((B, Lk, H, D), 2), _rand((B, Lk, H, D), 3)
    out = _run(q, k, v)
    ref = R.sdpa_blhd(q.float(), k.float(), v.float())
    err = R.max_rel_err(out.float().cpu(), ref)
    assert err < TOL, err


def test_fwd_lse():
    import b200vt.ops as ops
    q, k, v = _rand((1, 300, 2, 128), 4), _rand((1, 500, 2, 128), 5), _rand((1, 500, 2, 128), 6)
    scale = 1 / math.sqrt(128)
    o, lse = ops.attn_fwd(q.cuda(), k.cuda(), v.cuda(), None, None, None, 300, 500, scale)
    s = torch.einsum("bihd,bjhd->bhij", q.float(), k.float()) * scale
    assert torch.allclose(lse.cpu(), torch.logsumexp(s, dim=-1), atol=2e-3, rtol=1e-4)


def test_fwd_large_scores_lazy_rescale():
    # scores with a strongly growing running max exercise the rescale path
    q, k, v = _rand((1, 256, 1, 128), 7), _rand((1, 1024, 1, 128), 8), _rand((1, 1024, 1, 128), 9)
    k = (k.float() * torch.linspace(0.2, 6.0, 1024)[None, :, None, None]).to(torch.bfloat16)
    out = _run(q, k, v)
    ref = R.sdpa_blhd(q.float(), k.float(), v.float())
    assert R.max_rel_err(out.float().cpu(), ref) < TOL


def test_fwd_strided_views_of_fused_qkv():
    # hunyuan: q,k,v are views of one (B, L, 3, H, D) projection (models.py:165-166)
    B, L, H, D = 1, 300, 2, 128
    qkv = _rand((B, L, 3, H, D), 10).cuda()
    q, k, v = qkv.unbind(2)
    assert not q.is_contiguous()
    out = _run(q, k, v)
    ref = R.sdpa_blhd(q.float().cpu(), k.float().cpu(), v.float().cpu())
    assert R.max_rel_err(out.float().cpu(), ref) < TOL


def test_fwd_k_lens():
    # wan flash_attention(k_lens=...) (attention.py:62-71)
    q, k, v = _rand((2, 200, 2, 128), 11), _rand((2, 300, 2, 128), 12), _rand((2, 300, 2, 128), 13)
    k_lens = torch.tensor([300, 131], dtype=torch.int32)
    out = _run(q, k, v, k_lens=k_lens.cuda())
    ref = R.wan_flash_attention(q.float(), k.float(), v.float(), k_lens=k_lens)
    assert R.max_rel_err(out.float().cpu(), ref) < TOL


def test_fwd_varlen_two_segments_golden():
    # hunyuan mode="flash": cu_seqlens = [0, img+valid, img+max] per sample (attenion.py:34-57,108-119)
    g = load_golden("hunyuan_attention")
    import b200vt.functional as Fn
    q, k, v = g["q"].cuda(), g["k"].cuda(), g["v"].cuda()
    cu = R.hunyuan_cu_seqlens(g["text_mask"], g["img_len"]).cuda()
    out = Fn.hunyuan_attention(q, k, v, mode="flash", cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=150,
                               max_seqlen_kv=150, batch_size=2)
    assert out.shape == g["out_mask"].shape
    assert R.max_rel_err(out.float().cpu(), g["out_mask"]) < TOL
    out2 = Fn.hunyuan_attention(q, k, v, mode="torch")
    assert R.max_rel_err(out2.float().cpu(), g["out_plain"]) < TOL


def test_fwd_wan_golden():
    g = load_golden("wan_ops")
    import b200vt.functional as Fn
    out = Fn.wan_flash_attention(g["q"].cuda(), g["k"].cuda(), g["v"].cuda())
    assert out.dtype == torch.bfloat16
    assert R.max_rel_err(out.float().cpu(), g["sdpa_out"].float()) < TOL


def test_fwd_zero_length_keys():
    q, k, v = _rand((2, 130, 1, 64), 14), _rand((2, 64, 1, 64), 15), _rand((2, 64, 1, 64), 16)
    k_lens = torch.tensor([0, 64], dtype=torch.int32)
    out = _run(q, k, v, k_lens=k_lens.cuda()).float().cpu()
    assert float(out[0].abs().max()) == 0.0
    ref = R.sdpa_blhd(q[1:].float(), k[1:].float(), v[1:].float())
    assert R.max_rel_err(out[1:], ref) < TOL


def test_linearity_in_v_full_size_property():
    # size-independent property at a long sequence: attention is linear in V
    q, k = _rand((1, 4096, 2, 128), 17), _rand((1, 4096, 2, 128), 18)
    v1, v2 = _rand((1, 4096, 2, 128), 19), _rand((1, 4096, 2, 128), 20)
    o1, o2 = _run(q, k, v1).float(), _run(q, k, v2).float()
    o12 = _run(q, k, (v1.float() + v2.float()).to(torch.bfloat16)).float()
    assert R.max_rel_err(o12.cpu(), (o1 + o2).cpu()) < TOL
    # uniform keys -> output is the mean of V
    kz = torch.zeros_like(k)
    om = _run(q, kz, v1).float().cpu()
    assert R.max_rel_err(om, v1.float().mean(dim=1, keepdim=True).expand_as(om)) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------------------------------
def _grads_ref(q, k, v, do, mask=None, scale=None):
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    out = R.sdpa_blhd(qr, kr, vr, mask, scale)
    out.backward(do.float())
    return out.detach(), qr.grad, kr.grad, vr.grad


def _grads_cuda(q, k, v, do, **kw):
    import b200vt.functional as Fn
    qc, kc, vc = (t.cuda().clone().requires_grad_(True) for t in (q, k, v))
    out = Fn.attention_blhd(qc, kc, vc, **kw)
    out.backward(do.cuda())
    torch.cuda.synchronize()
    return out.detach(), qc.grad, kc.grad, vc.grad


def _check_grads(got, ref, names=("out", "dq", "dk", "dv")):
    for n, a, b in zip(names, got, ref):
        a = a.float().cpu()
        cos, err = R.cosine(a, b), R.max_rel_err(a, b)
        assert cos > 0.999, (n, cos, err)
        assert err < 3e-2, (n, cos, err)


@pytest.mark.parametrize("B,Lq,Lk,H,D", [
    (1, 128, 128, 1, 128), (1, 128, 128, 1, 64), (1, 256, 384, 2, 128), (2, 200, 333, 2, 128), (1, 1000, 77, 3, 64),
    (2, 40, 40, 5, 64), (1, 640, 640, 2, 64), (1, 513, 1025, 1, 128), (1, 2560, 2560, 2, 64), (1, 1500, 1500, 2, 128),
])
def test_bwd_matches_oracle(B, Lq, Lk, H, D):
    q, k, v = _rand((B, Lq, H, D), 21), _rand((B, Lk, H, D), 22), _rand((B, Lk, H, D), 23)
    do = _rand((B, Lq, H, D), 24)
    _check_grads(_grads_cuda(q, k, v, do), _grads_ref(q, k, v, do))


def test_bwd_k_lens():
    q, k, v = _rand((2, 200, 2, 128), 25), _rand((2, 300, 2, 128), 26), _rand((2, 300, 2, 128), 27)
    do = _rand((2, 200, 2, 128), 28)
    k_lens = torch.tensor([300, 131], dtype=torch.int32)
    mask = (torch.arange(300)[None, :] < k_lens[:, None])[:, None, None, :]
    got = _grads_cuda(q, k, v, do, k_lens=k_lens.cuda())
    ref = _grads_ref(q, k, v, do, mask)
    _check_grads(got, ref)
    assert float(got[2][1, 131:].abs().max()) == 0.0 and float(got[3][1, 131:].abs().max()) == 0.0


def test_bwd_varlen_two_segments():
    g = load_golden("hunyuan_attention")
    q, k, v = g["q"][:1], g["k"][:1], g["v"][:1]  # one sample: segments [0,140) and [140,150)
    do = _rand(tuple(q.shape), 29)
    cu = torch.tensor([0, 140, 150], dtype=torch.int32)
    mask = R.varlen_block_mask(cu.tolist(), 150)[None, None]
    got = _grads_cuda(q, k, v, do, cu_seqlens_q=cu.cuda(), cu_seqlens_k=cu.cuda(), max_seqlen_q=150, max_seqlen_k=150)
    _check_grads(got, _grads_ref(q.float(), k.float(), v.float(), do, mask))


def test_bwd_strided_fused_qkv_and_softmax_scale():
    qkv = _rand((1, 300, 3, 2, 64), 30)
    q, k, v = qkv.unbind(2)
    do = _rand((1, 300, 2, 64), 31)
    import b200vt.functional as Fn
    qkv_c = qkv.cuda().clone().requires_grad_(True)
    qc, kc, vc = qkv_c.unbind(2)
    out = Fn.attention_blhd(qc, kc, vc, softmax_scale=0.2)
    out.backward(do.cuda())
    ref = _grads_ref(q, k, v, do, scale=0.2)
    gq, gk, gv = qkv_c.grad.unbind(2)
    _check_grads((out.detach(), gq, gk, gv), ref)


def test_bwd_gradient_identities_long_sequence():
    # size-independent properties at a long sequence: sum_j dS_ij = 0 => dq is orthogonal to nothing in general, but
    # (1) dv = P^T dO implies sum over keys of dv equals sum over queries of dO (rows of P sum to 1);
    # (2) the loss sum(o * do) is invariant to adding a constant vector c to all keys: dk must sum to ~0 along keys... per head dim weighted by q.
    q, k, v = _rand((1, 8192, 1, 128), 32), _rand((1, 8192, 1, 128), 33), _rand((1, 8192, 1, 128), 34)
    do = _rand((1, 8192, 1, 128), 35)
    out, dq, dk, dv = _grads_cuda(q, k, v, do)
    lhs, rhs = dv.float().sum(dim=1).cpu(), do.float().sum(dim=1)
    assert R.max_rel_err(lhs, rhs) < 2e-2
    # shifting every key by the same vector leaves softmax unchanged -> sum_j dk_j = 0 (up to bf16 rounding of dk)
    assert float(dk.float().sum(dim=1).abs().max()) < 2e-2 * float(dk.float().abs().sum(dim=1).max())


def test_host_buffer_entry_point_matches_device_path():
    # functional.HostAttention: pinned host tensors, head groups pipelined over three streams
    import b200vt.functional as Fn
    B, L, H, D = 2, 333, 4, 64
    q, k, v, do = (_rand((B, L, H, D), s).pin_memory() for s in (21, 22, 23, 24))
    outs = [torch.empty((B, L, H, D), dtype=torch.bfloat16).pin_memory() for _ in range(4)]
    ha = Fn.HostAttention(B, L, H, D, head_groups=2)
    for _ in range(3):  # consecutive calls reuse the device buffers and pipeline into one another
        ha(q, k, v, do, *outs)
    ha.synchronize()
    torch.cuda.synchronize()
    qd, kd, vd = (t.cuda().requires_grad_(True) for t in (q, k, v))
    ref = Fn.attention_blhd(qd, kd, vd)
    gq, gk, gv = torch.autograd.grad(ref, (qd, kd, vd), do.cuda())
    for name, got, want in zip(("o", "dq", "dk", "dv"), outs, (ref, gq, gk, gv)):
        # same kernels on the same values; dq is accumulated with fp32 atomics whose order is not fixed
        err = R.max_rel_err(got.float(), want.detach().float().cpu())
        assert err < (1e-2 if name == "dq" else 1e-6), (name, err)


def test_host_ulysses_entry_point_single_rank_pipelines_across_calls():
    """sp.HostUlyssesAttention at SP world size 1 (no process group): pinned host shards in, pinned results out, head groups
    and CONSECUTIVE CALLS pipelined; equals joint attention of [img; txt] + autograd on device tensors."""
    import b200vt.functional as Fn
    import b200vt.sp as sp
    S, T, H, D = 300, 20, 4, 128
    q, k, v = (_rand((1, S, H, D), s).pin_memory() for s in (41, 42, 43))
    tq, tk, tv = (_rand((1, T, H, D), s).pin_memory() for s in (44, 45, 46))
    do = _rand((1, S + T, H, D), 47).pin_memory()
    outs = [[torch.empty(shape, dtype=torch.bfloat16).pin_memory() for shape in
             ((1, S + T, H, D), (1, S, H, D), (1, S, H, D), (1, S, H, D), (1, T, H, D), (1, T, H, D), (1, T, H, D))]
            for _ in range(3)]
    ha = sp.HostUlyssesAttention(S, T, H, D, head_groups=2)
    for o in outs:  # three calls back to back, no synchronisation in between
        ha(q, k, v, do, o[0], o[1], o[2], o[3], tq, tk, tv, o[4], o[5], o[6])
    ha.synchronize()
    torch.cuda.synchronize()
    leaves = [t.cuda().requires_grad_(True) for t in (q, k, v, tq, tk, tv)]
    ref = Fn.attention_blhd(*(torch.cat([leaves[i], leaves[i + 3]], 1) for i in range(3)))
    grads = torch.autograd.grad(ref, leaves, do.cuda())
    want = (ref, *grads)
    for o in outs:
        for name, got, w in zip(("out", "dq", "dk", "dv", "dtq", "dtk", "dtv"), o, want):
            err = R.max_rel_err(got.float(), w.detach().float().cpu())
            assert err < (1e-2 if name in ("dq", "dtq") else 1e-6), (name, err)
