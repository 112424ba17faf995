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
