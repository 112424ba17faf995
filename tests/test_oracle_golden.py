"""Pin the oracle (oracle/ref_ops.py) against fixtures recorded from the reference's own code
(tests/golden/make_golden.py ran /root/reference on CPU; inputs and weights are bf16-representable fp32)."""
import torch

from conftest import load_golden
from oracle import ref_ops as R

TOL = 2e-5  # fp32 vs fp32, different summation order


def f32(t):
    return t.float()


def test_hunyuan_attention_torch_mode():
    g = load_golden("hunyuan_attention")
    q, k, v = f32(g["q"]), f32(g["k"]), f32(g["v"])
    assert R.max_rel_err(R.hunyuan_attention_torch(q, k, v), g["out_plain"]) < TOL
    assert R.max_rel_err(R.hunyuan_attention_torch(q, k, v, g["attn_mask"]), g["out_mask"]) < TOL


def test_hunyuan_attention_torch_fused_form_matches_golden():
    # the F.scaled_dot_product_attention form that bench.py times on the host cores
    g = load_golden("hunyuan_attention")
    q, k, v = f32(g["q"]), f32(g["k"]), f32(g["v"])
    assert R.max_rel_err(R.hunyuan_attention_torch_fused(q, k, v), g["out_plain"]) < TOL
    assert R.max_rel_err(R.hunyuan_attention_torch_fused(q, k, v, g["attn_mask"]), g["out_mask"]) < TOL


def test_hunyuan_flash_semantics_equal_masked_torch_mode():
    g = load_golden("hunyuan_attention")
    q, k, v = f32(g["q"]), f32(g["k"]), f32(g["v"])
    cu = R.hunyuan_cu_seqlens(g["text_mask"], g["img_len"])
    assert cu.tolist() == [0, 140, 150, 300, 300]
    out = R.hunyuan_attention_flash_semantics(q, k, v, cu)
    assert R.max_rel_err(out, g["out_mask"]) < TOL


def test_hunyuan_rmsnorm_rope_modulate_gate():
    g = load_golden("hunyuan_norm_rope")
    w = f32(g["w"])
    xq, xk = f32(g["xq"]), f32(g["xk"])
    assert R.max_rel_err(R.hunyuan_rmsnorm(xq, w, g["eps"]), g["norm_q"]) < TOL
    cos, sin = R.hunyuan_nd_rope(g["rope_dim_list"], g["rope_sizes"], theta=g["theta"])
    assert torch.allclose(cos, g["cos"].float(), atol=1e-6) and torch.allclose(sin, g["sin"].float(), atol=1e-6)
    assert R.max_rel_err(R.hunyuan_apply_rotary_emb(R.hunyuan_rmsnorm(xq, w, g["eps"]), cos, sin), g["rope_q"]) < TOL
    assert R.max_rel_err(R.hunyuan_apply_rotary_emb(R.hunyuan_rmsnorm(xk, w, g["eps"]), cos, sin), g["rope_k"]) < TOL
    xm = f32(g["xm"])
    assert R.max_rel_err(R.modulate(xm, f32(g["shift"]), f32(g["scale"])), g["modulated"]) < TOL
    assert R.max_rel_err(R.apply_gate(xm, f32(g["gate"])), g["gated"]) < TOL


def test_wan_ops():
    g = load_golden("wan_ops")
    q, k, v = f32(g["q"]), f32(g["k"]), f32(g["v"])
    # the reference's SDPA fallback runs in bf16 (attention.py:171-175 casts to dtype=bfloat16): bf16-level tolerance
    assert R.max_rel_err(R.wan_flash_attention(q, k, v), g["sdpa_out"]) < 2e-2
    assert R.max_rel_err(R.wan_attention_sdpa_fallback(q.bfloat16(), k.bfloat16(), v.bfloat16()), g["sdpa_out"]) < 1e-2
    freqs = R.wan_freqs_table(128)
    assert R.max_rel_err(R.wan_rope_apply(q, g["grid"], freqs), g["roped"]) < TOL
    # the cos/sin-table form used by the fused kernel is the same rotation
    cos, sin = R.wan_rope_cos_sin(g["grid"][0].tolist(), freqs)
    n0 = int(g["grid"][0].prod())
    roped0 = R.hunyuan_apply_rotary_emb(q[:1, :n0], cos, sin)
    assert R.max_rel_err(roped0, g["roped"][:1, :n0]) < TOL
    assert torch.equal(g["roped"][0, n0:], q[0, n0:])  # tokens past f*h*w are passed through (model.py:62)
    assert R.max_rel_err(R.wan_rmsnorm(f32(g["xr"]), f32(g["rms_w"])), g["rms_out"]) < TOL
    assert R.max_rel_err(R.layer_norm(f32(g["xr"]), None, None, 1e-6), g["ln_out"]) < TOL


def test_lvdm_groupnorm_silu():
    g = load_golden("lvdm_groupnorm")
    x = f32(g["x"])
    assert R.max_rel_err(R.groupnorm_silu(x, f32(g["weight"]), f32(g["bias"]), 32, g["eps"], False), g["out"]) < TOL
    assert R.max_rel_err(R.groupnorm_silu(x, f32(g["weight"]), f32(g["bias"]), 32, g["eps"], True), g["out_silu"]) < TOL


def _lvdm_cross_attention(case):
    """CrossAttention.forward (attention.py:101-170) restated with the oracle's pieces."""
    sd = {k: v.float() for k, v in case["sd"].items()}
    kw = case["kw"]
    h = kw["heads"]
    x = case["x"].float()
    ctx = x if case["context"] is None else case["context"].float()
    scale = kw["dim_head"] ** -0.5
    q = x @ sd["to_q.weight"].T
    k_ip = v_ip = None
    if case["context"] is not None:
        if kw.get("img_cross_attention", False):
            ctx, ctx_img = ctx[:, :77], ctx[:, 77:]
            k_ip, v_ip = ctx_img @ sd["to_k_ip.weight"].T, ctx_img @ sd["to_v_ip.weight"].T
        else:
            ctx = ctx[:, :77]
    k, v = ctx @ sd["to_k.weight"].T, ctx @ sd["to_v.weight"].T
    qh, kh, vh = (R.lvdm_split_heads(t, h) for t in (q, k, v))
    rel_k = rel_v = None
    if kw.get("relative_position", False):
        T = kw["temporal_length"]
        rel_k = R.lvdm_relative_position(sd["relative_position_k.embeddings_table"], qh.shape[1], kh.shape[1], T)
        rel_v = R.lvdm_relative_position(sd["relative_position_v.embeddings_table"], qh.shape[1], vh.shape[1], T)
    mask = case["mask"]
    out = R.lvdm_merge_heads(R.lvdm_attention_core(qh, kh, vh, scale, rel_k, rel_v, mask), h)
    if k_ip is not None:
        out_ip = R.lvdm_merge_heads(
            R.lvdm_attention_core(qh, R.lvdm_split_heads(k_ip, h), R.lvdm_split_heads(v_ip, h), scale), h)
        out = out + kw["img_cross_attention_scale"] * out_ip
    return out @ sd["to_out.0.weight"].T + sd["to_out.0.bias"]


def test_lvdm_cross_attention_all_variants():
    cases = load_golden("lvdm_cross_attention")
    assert set(cases) == {"self", "cross", "temporal_relpos", "temporal_relpos_causal", "img_cross"}
    for name, case in cases.items():
        err = R.max_rel_err(_lvdm_cross_attention(case), case["out"])
        assert err < 5e-5, (name, err)


def test_chunked_attention_reference_equals_explicit_softmax_and_autograd():
    # oracle.attention_fwd_bwd_chunked (used at the BASELINE sizes on the GPU box) against sdpa_blhd + autograd, with
    # the two-segment varlen mask of the golden fixture, a key-length mask, ragged chunking, and rows outside every segment
    g = load_golden("hunyuan_attention")
    q, k, v = (f32(g[n])[0, :, 1].contiguous() for n in ("q", "k", "v"))  # one sample, one head: (150, D)
    gen = torch.Generator().manual_seed(5)
    do = torch.randn(q.shape, generator=gen)
    for segments, k_len in (([0, 140, 150], None), (None, 131), (None, None), ([0, 100, 120], None)):
        n = 150 if segments is None else segments[-1]  # rows / keys past the last segment take no part
        qr, kr, vr = (t[:n].clone().requires_grad_(True) for t in (q, k, v))
        mask = None
        if segments is not None:
            mask = R.varlen_block_mask(segments, n)[None, None]
        if k_len is not None:
            mask = (torch.arange(n) < k_len)[None, None, None, :]
        out = R.sdpa_blhd(qr[None, :, None], kr[None, :, None], vr[None, :, None], mask)[0, :, 0]
        out.backward(do[:n])
        o, lse, dq, dk, dv = R.attention_fwd_bwd_chunked(q, k, v, do, segments=segments, k_len=k_len, chunk=64)
        assert R.max_rel_err(o[:n], out.detach()) < TOL
        for got, want in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
            assert R.max_rel_err(got[:n], want) < TOL
            assert n == 150 or float(got[n:].abs().max()) == 0.0
        assert n == 150 or float(o[n:].abs().max()) == 0.0


def test_kernel_side_approximations_are_within_their_stated_bounds():
    """The closed forms the CUDA kernels evaluate inside exact formulas (oracle restatements in float32 numpy): erf by
    Abramowitz-Stegun 7.1.26 (GEGLU: exact-erf GELU and its derivative) and SiLU / SiLU' through one tanh (GroupNorm+SiLU),
    against float64 references over the whole useful range."""
    import math

    import numpy as np
    from oracle import ref_ops as R
    x = np.linspace(-6.0, 6.0, 200001)
    exact = np.array([math.erf(v) for v in x])
    # published bound 1.5e-7 in exact arithmetic; evaluated in float32 (coefficients, exp, the final 1 - ...) 5.3e-7
    assert np.abs(R.erf_abramowitz_stegun_f32(x).astype(np.float64) - exact).max() <= 6e-7
    g = np.linspace(-12.0, 12.0, 200001)
    cdf = 0.5 * (1.0 + np.array([math.erf(v / math.sqrt(2.0)) for v in g]))
    pdf = np.exp(-0.5 * g * g) / math.sqrt(2.0 * math.pi)
    y, dy = R.gelu_and_grad_f32(g)
    assert np.abs(y.astype(np.float64) - g * cdf).max() <= 2e-6      # far below one bf16 ulp of the output
    assert np.abs(dy.astype(np.float64) - (cdf + g * pdf)).max() <= 2e-6
    z = np.linspace(-30.0, 30.0, 200001)
    sig = 1.0 / (1.0 + np.exp(-z))
    s, ds = R.silu_and_grad_by_tanh_f32(z)
    assert np.abs(s.astype(np.float64) - z * sig).max() <= 4e-6
    assert np.abs(ds.astype(np.float64) - sig * (1.0 + z * (1.0 - sig))).max() <= 2e-6


def test_fma_pipe_exp2_polynomial_error_bound():
    """csrc/sm100_ptx.cuh ex2_poly2 (restated bit for bit in float32 numpy): relative error against 2^x over the softmax's
    argument range (x <= 0 after the running maximum is subtracted; down to the -125 clamp), and the clamp itself."""
    import numpy as np
    from oracle import ref_ops as R
    x = np.concatenate([np.linspace(-124.9, 0.0, 400001), np.linspace(0.0, 20.0, 50001)])
    got = R.ex2_poly_f32(x).astype(np.float64)
    want = np.exp2(x.astype(np.float32).astype(np.float64))
    assert np.abs(got / want - 1.0).max() <= 7.5e-5
    assert float(R.ex2_poly_f32(np.array([-np.inf]))[0]) == float(R.ex2_poly_f32(np.array([-125.0]))[0]) < 3e-38
    assert abs(float(R.ex2_poly_f32(np.array([0.0]))[0]) - 1.0) <= 7.5e-5
