"""lvdm drop-in body and the temporal micro-attention kernel on the GPU, against fixtures recorded from the reference
(tests/golden/lvdm_*.pt) and against the oracle. Tolerances: BASELINE.md §5."""
import pytest
import torch

from conftest import load_golden
from helpers import CrossAttentionShell
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _run_case(case):
    import b200vt.functional as Fn
    dev = torch.device("cuda")
    m = CrossAttentionShell.from_fixture(case, dev)
    x = case["x"].to(dev, torch.bfloat16)
    ctx = None if case["context"] is None else case["context"].to(dev, torch.bfloat16)
    mask = None if case["mask"] is None else case["mask"].to(dev).expand(x.shape[0], -1, -1)
    return m, x, Fn.lvdm_cross_attention_forward(m, x, context=ctx, mask=mask)


@pytest.mark.parametrize("name", ["self", "cross", "img_cross"])
def test_cross_attention_forward_matches_reference_fixture(name):
    case = load_golden("lvdm_cross_attention")[name]
    _, _, out = _run_case(case)
    assert R.max_rel_err(out.float().cpu(), case["out"].float()) < TOL


@pytest.mark.parametrize("name", ["temporal", "temporal_causal"])
def test_temporal_attention_module_matches_reference_fixture(name):
    case = load_golden("lvdm_temporal_attention")[name]
    _, _, out = _run_case(case)
    assert R.max_rel_err(out.float().cpu(), case["out"].float()) < TOL


def test_temporal_causal_backward_matches_reference_fixture():
    import b200vt.functional as Fn
    case = load_golden("lvdm_temporal_attention")["temporal_causal"]
    dev = torch.device("cuda")
    m = CrossAttentionShell.from_fixture(case, dev)
    x = case["x"].to(dev, torch.bfloat16).requires_grad_(True)
    mask = case["mask"].to(dev).expand(x.shape[0], -1, -1)
    out = Fn.lvdm_cross_attention_forward(m, x, mask=mask)
    out.backward(case["dout"].to(dev, torch.bfloat16))
    assert R.cosine(x.grad.float().cpu(), case["dx"].float()) > 0.999
    assert R.cosine(m.to_q.weight.grad.float().cpu(), case["dw_q"].float()) > 0.999
    assert R.cosine(m.to_v.weight.grad.float().cpu(), case["dw_v"].float()) > 0.999


@pytest.mark.parametrize("name", ["temporal_relpos", "temporal_relpos_causal"])
def test_relative_position_attention_forward_matches_reference_fixture(name):
    """VideoCrafter1's relative-position temporal attention (attention.py:19-42, 129-133, 145-148) on the CUDA path."""
    case = load_golden("lvdm_cross_attention")[name]
    _, _, out = _run_case(case)
    assert R.max_rel_err(out.float().cpu(), case["out"].float()) < TOL


@pytest.mark.parametrize("tag", ["plain", "causal"])
def test_relative_position_attention_backward_matches_reference_fixture(tag):
    import b200vt.functional as Fn
    g = load_golden("lvdm_extra")["relpos"]
    dev = torch.device("cuda")
    m = CrossAttentionShell.from_fixture(g, dev)
    x = g["x"].to(dev, torch.bfloat16).requires_grad_(True)
    mask = g["mask"].to(dev).expand(x.shape[0], -1, -1) if tag == "causal" else None
    out = Fn.lvdm_cross_attention_forward(m, x, mask=mask)
    r = g[tag]
    assert R.max_rel_err(out.float().cpu(), r["out"].float()) < TOL
    out.backward(g["d_out"].to(dev, torch.bfloat16))
    assert R.cosine(x.grad.float().cpu(), r["d_x"].float()) > 0.999
    assert R.cosine(m.to_q.weight.grad.float().cpu(), r["d_to_q"].float()) > 0.999
    assert R.cosine(m.relative_position_k.embeddings_table.grad.float().cpu(), r["d_rel_k"].float()) > 0.995
    assert R.cosine(m.relative_position_v.embeddings_table.grad.float().cpu(), r["d_rel_v"].float()) > 0.995


def test_fp32_activations_stay_on_reference_path():
    import b200vt.functional as Fn
    case = load_golden("lvdm_cross_attention")["self"]
    dev = torch.device("cuda")
    m = CrossAttentionShell.from_fixture(case, dev, dtype=torch.float32)
    with pytest.raises(Fn.Unsupported):
        Fn.lvdm_cross_attention_forward(m, case["x"].to(dev, torch.float32))


def _rand(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16)


@pytest.mark.parametrize("B,N,H,D,masked", [(3, 16, 5, 64, False), (7, 16, 2, 64, True), (2, 1, 3, 64, False),
                                             (5, 7, 2, 64, True), (4, 32, 2, 64, False), (3, 16, 2, 128, True),
                                             (1000, 16, 5, 64, False)])
def test_temporal_kernel_fwd_bwd_vs_oracle(B, N, H, D, masked):
    import b200vt.functional as Fn
    q, k, v, do = (_rand((B, N, H, D), s) for s in (1, 2, 3, 4))
    mask = torch.tril(torch.ones(N, N)) if masked else None
    scale = D ** -0.5
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))

    def heads_first(t):  # (B,N,H,D) -> (B*H, N, D)
        return t.permute(0, 2, 1, 3).reshape(B * H, N, D)

    ref = R.lvdm_attention_core(heads_first(qr), heads_first(kr), heads_first(vr), scale,
                                mask=None if mask is None else mask[None])
    ref = ref.view(B, H, N, D).permute(0, 2, 1, 3)
    ref.backward(do.float())
    qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q, k, v))
    out = Fn.temporal_attention(qc, kc, vc, softmax_scale=scale, mask=None if mask is None else mask.cuda())
    out.backward(do.cuda())
    assert R.max_rel_err(out.float().cpu(), ref.detach()) < TOL
    for got, want in ((qc.grad, qr.grad), (kc.grad, kr.grad), (vc.grad, vr.grad)):
        if float(want.abs().max()) < 1e-6:  # N == 1: softmax over one key is constant, dq = dk = 0 exactly
            assert float(got.float().abs().max()) < 1e-3
            continue
        assert R.cosine(got.float().cpu(), want) > 0.999
        assert R.max_rel_err(got.float().cpu(), want) < TOL


def test_temporal_kernel_strided_views():
    # q, k, v as views of one fused (B, N, 3, H, D) projection
    import b200vt.functional as Fn
    B, N, H, D = 9, 16, 4, 64
    qkv = _rand((B, N, 3, H, D), 11).cuda()
    q, k, v = qkv.unbind(2)
    out = Fn.temporal_attention(q, k, v)
    ref = R.sdpa_blhd(q.float().cpu(), k.float().cpu(), v.float().cpu())
    assert R.max_rel_err(out.float().cpu(), ref) < TOL


@pytest.mark.parametrize("kind", ["self", "cross", "temporal"])
def test_lora_adapters_on_the_projections_dense_gemm_path(kind, monkeypatch):
    """peft-style LoRA on to_q / to_k / to_v (lvdm/ddpm3d.py:112-117, vc2_t2v_lora.yaml): the drop-in evaluates them as
    base GEMM + dense-delta GEMM (functional.lora_merged_projections). Output, input gradient and the adapter gradients
    must match the modules' own forward (B200VT_LORA_MERGE=0 route: same kernels, per-module projections) and the fp32 oracle."""
    import b200vt.functional as Fn
    from helpers import PeftLikeLinear
    dev = torch.device("cuda")
    torch.manual_seed(11)
    dim, heads, d = 320, 5, 64
    b, n = (8, 16) if kind == "temporal" else (2, 640)
    m = CrossAttentionShell(dim, context_dim=None if kind != "cross" else 1024, heads=heads, dim_head=d)
    m.to_q, m.to_k, m.to_v = PeftLikeLinear(m.to_q), PeftLikeLinear(m.to_k), PeftLikeLinear(m.to_v)
    m = m.to(dev, torch.bfloat16)
    x = torch.randn(b, n, dim, device=dev, dtype=torch.bfloat16)
    ctx = torch.randn(b, 77, 1024, device=dev, dtype=torch.bfloat16) if kind == "cross" else None
    gy = torch.randn(b, n, dim, device=dev, dtype=torch.bfloat16)

    def run(merge):
        monkeypatch.setattr(Fn, "_LORA_MERGE", merge)
        xr = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        y = Fn.lvdm_cross_attention_forward(m, xr, context=ctx)
        y.backward(gy)
        return y, xr.grad, {k: p.grad.float().clone() for k, p in m.named_parameters() if p.grad is not None}

    y1, dx1, g1 = run(True)
    y0, dx0, g0 = run(False)
    assert set(g1) == set(g0) and any("lora_A" in k for k in g1) and not any("base_layer" in k for k in g1)
    assert R.max_rel_err(y1, y0) <= TOL and R.max_rel_err(dx1, dx0) <= TOL
    for k in g0:
        assert R.cosine(g1[k], g0[k]) >= 0.999, k
    # fp32 oracle of the whole module (adapters included)
    mf = CrossAttentionShell(dim, context_dim=None if kind != "cross" else 1024, heads=heads, dim_head=d)
    mf.to_q, mf.to_k, mf.to_v = PeftLikeLinear(mf.to_q), PeftLikeLinear(mf.to_k), PeftLikeLinear(mf.to_v)
    mf.load_state_dict({k: v.float() for k, v in m.state_dict().items()})
    mf = mf.to(dev)
    xf = x.float().requires_grad_(True)
    cf = xf if ctx is None else ctx.float()
    q, k, v = mf.to_q(xf), mf.to_k(cf), mf.to_v(cf)
    o = R.sdpa_blhd(q.view(b, n, heads, d), k.view(b, -1, heads, d), v.view(b, -1, heads, d))
    yf = mf.to_out(o.reshape(b, n, heads * d))
    yf.backward(gy.float())
    assert R.max_rel_err(y1, yf) <= TOL and R.cosine(dx1, xf.grad) >= 0.999
    for k, p in mf.named_parameters():
        if p.grad is not None and k in g1:
            assert R.cosine(g1[k], p.grad) >= 0.999, k
