"""Generate tests/golden/*.pt by running the UNMODIFIED reference code (/root/reference) on seeded CPU inputs.

Run in the development container only (`python tests/golden/make_golden.py`; `--check` regenerates everything in
memory and diffs it bit for bit against the committed fixtures); the GPU box has no /root/reference and only reads the
committed fixtures. Import shims cover wheels that are absent here: colorama,
omegaconf, diffusers — none of them takes part in the arithmetic being recorded.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 20230211  # the reference's default --seed (scripts/train_new.py:34)


def install_shims():
    sys.path.insert(0, REF)
    col = types.ModuleType("colorama")
    col.Fore = types.SimpleNamespace(**{k: "" for k in ("RED", "GREEN", "YELLOW", "BLUE", "CYAN", "MAGENTA", "WHITE", "RESET")})
    col.Style = types.SimpleNamespace(RESET_ALL="", BRIGHT="", DIM="", NORMAL="")
    col.init = lambda *a, **k: None
    sys.modules.setdefault("colorama", col)
    om = types.ModuleType("omegaconf")
    om.DictConfig = dict
    om.ListConfig = list
    om.OmegaConf = type("OmegaConf", (), {})
    sys.modules.setdefault("omegaconf", om)
    dif = types.ModuleType("diffusers")
    difm = types.ModuleType("diffusers.models")
    difm.ModelMixin = torch.nn.Module
    difc = types.ModuleType("diffusers.configuration_utils")
    difc.ConfigMixin = object
    difc.register_to_config = lambda f: f
    difm.__path__ = []
    difmu = types.ModuleType("diffusers.models.modeling_utils")
    difmu.ModelMixin = torch.nn.Module
    dif.models, dif.configuration_utils, difm.modeling_utils = difm, difc, difmu
    sys.modules.setdefault("diffusers", dif)
    sys.modules.setdefault("diffusers.models", difm)
    sys.modules.setdefault("diffusers.models.modeling_utils", difmu)
    sys.modules.setdefault("diffusers.configuration_utils", difc)
    # hyvideo_i2v/utils/helpers.py:11 imports deepspeed at module level; nothing on the block path uses it
    sys.modules.setdefault("deepspeed", types.ModuleType("deepspeed"))


def dezero(module: torch.nn.Module, gen: torch.Generator):
    """Re-draw every all-zero parameter (zero_module / ModulateDiT / FinalLayer init) so parity is not vacuous."""
    with torch.no_grad():
        for p in module.parameters():
            if p.numel() > 0 and float(p.abs().max()) == 0.0:
                p.copy_(RN(p.shape, generator=gen) * 0.02)


def _compact(obj):
    """Store tensors whose values are exactly bf16-representable (inputs, weights) as bf16; keep the rest (outputs)."""
    if isinstance(obj, torch.Tensor):
        if obj.dtype == torch.float32 and torch.equal(obj, obj.to(torch.bfloat16).to(torch.float32)):
            return obj.to(torch.bfloat16)
        return obj.clone()
    if isinstance(obj, dict):
        return {k: _compact(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_compact(v) for v in obj)
    return obj


CHECK = False      # --check: compare with the committed fixtures instead of writing them
MISMATCHES = []


def _diff(a, b, path, out):
    if isinstance(a, torch.Tensor):
        if not (isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)):
            out.append(path)
    elif isinstance(a, dict):
        if not isinstance(b, dict) or set(a) != set(b):
            out.append(path + " (keys)")
        else:
            for k in a:
                _diff(a[k], b[k], f"{path}.{k}", out)
    elif isinstance(a, (list, tuple)):
        if not isinstance(b, (list, tuple)) or len(a) != len(b):
            out.append(path + " (len)")
        else:
            for i, (x, y) in enumerate(zip(a, b)):
                _diff(x, y, f"{path}[{i}]", out)
    elif a != b:
        out.append(path)


def save(name, obj):
    path = os.path.join(OUT, name + ".pt")
    obj = _compact(obj)
    if CHECK:
        bad = []
        _diff(obj, torch.load(path, map_location="cpu"), name, bad)
        MISMATCHES.extend(bad)
        print(f"{name}: {'bit-identical' if not bad else 'DIFFERS at ' + ', '.join(bad[:6])}")
        return
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def r16(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16-representable values, keep fp32 storage: the reference then runs in fp32 on exactly the values a
    bf16 CUDA kernel will see, so fixtures carry no input-rounding error."""
    return t.to(torch.bfloat16).to(torch.float32)


def RN(*shape, generator=None):
    return r16(torch.randn(*shape, generator=generator))


def round_params(module: torch.nn.Module):
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(r16(p))
        for b in module.buffers():
            if b.is_floating_point():
                b.copy_(r16(b))
    return module


def sd(module):
    """state dict stored as bf16 (values are bf16-representable, see round_params) to keep fixtures small."""
    return {k: (v.detach().to(torch.bfloat16) if v.is_floating_point() else v.detach().clone())
            for k, v in module.state_dict().items()}


def _seeded(offset: int) -> torch.Generator:
    """Module constructors draw their initial weights from the GLOBAL generator: every fixture function re-seeds it and
    owns a private generator for its inputs, so each fixture regenerates bit-identically whatever ran before it."""
    torch.manual_seed(SEED + offset)
    return torch.Generator().manual_seed(SEED + offset)


def golden_lvdm():
    gen = _seeded(0)
    from videotuna.models.lvdm.modules import attention as A
    assert not A.XFORMERS_IS_AVAILBLE
    cases = {}
    # spatial self-attention (ragged N), cross-attention with 77-token truncation, temporal w/ relative position + mask
    m = round_params(A.CrossAttention(query_dim=128, context_dim=None, heads=2, dim_head=64))
    x = RN(2, 200, 128, generator=gen)
    cases["self"] = dict(kw=dict(query_dim=128, context_dim=None, heads=2, dim_head=64), sd=sd(m), x=x, context=None,
                         mask=None, out=m(x).detach())
    m = round_params(A.CrossAttention(query_dim=128, context_dim=96, heads=2, dim_head=64))
    ctx = RN(2, 90, 96, generator=gen)  # > 77: reference truncates (attention.py:120-121)
    cases["cross"] = dict(kw=dict(query_dim=128, context_dim=96, heads=2, dim_head=64), sd=sd(m), x=x, context=ctx,
                          mask=None, out=m(x, context=ctx).detach())
    m = A.CrossAttention(query_dim=128, context_dim=None, heads=2, dim_head=64, relative_position=True,
                         temporal_length=16)
    round_params(m)
    xt = RN(6, 16, 128, generator=gen)
    mask = torch.tril(torch.ones(1, 16, 16))
    cases["temporal_relpos"] = dict(kw=dict(query_dim=128, context_dim=None, heads=2, dim_head=64,
                                            relative_position=True, temporal_length=16), sd=sd(m), x=xt, context=None,
                                    mask=None, out=m(xt).detach())
    cases["temporal_relpos_causal"] = dict(kw=cases["temporal_relpos"]["kw"], sd=sd(m), x=xt, context=None, mask=mask,
                                           out=m(xt, mask=mask.expand(6, -1, -1)).detach())
    m = A.CrossAttention(query_dim=128, context_dim=96, heads=2, dim_head=64, img_cross_attention=True,
                         img_cross_attention_scale=0.7)
    round_params(m)
    ctx2 = RN(2, 77 + 16, 96, generator=gen)
    cases["img_cross"] = dict(kw=dict(query_dim=128, context_dim=96, heads=2, dim_head=64, img_cross_attention=True,
                                      img_cross_attention_scale=0.7), sd=sd(m), x=x, context=ctx2, mask=None,
                              out=m(x, context=ctx2).detach())
    save("lvdm_cross_attention", cases)

    # whole transformer wrappers (GroupNorm + permutes + LayerNorm + attention + GEGLU), de-zeroed
    st = A.SpatialTransformer(in_channels=128, n_heads=2, d_head=64, depth=1, context_dim=96, use_linear=True,
                              use_checkpoint=False)
    dezero(st, gen)
    round_params(st)
    xs = RN(3, 128, 5, 8, generator=gen)
    ctxs = RN(3, 77, 96, generator=gen)
    tt = A.TemporalTransformer(in_channels=128, n_heads=2, d_head=64, depth=1, use_linear=True, use_checkpoint=False,
                               only_self_att=True, temporal_length=16)
    dezero(tt, gen)
    round_params(tt)
    xtt = RN(1, 128, 16, 3, 4, generator=gen)
    save("lvdm_transformers", dict(
        spatial=dict(kw=dict(in_channels=128, n_heads=2, d_head=64, depth=1, context_dim=96, use_linear=True,
                             use_checkpoint=False), sd=sd(st), x=xs, context=ctxs, out=st(xs, ctxs).detach()),
        temporal=dict(kw=dict(in_channels=128, n_heads=2, d_head=64, depth=1, use_linear=True, use_checkpoint=False,
                              only_self_att=True, temporal_length=16), sd=sd(tt), x=xtt, out=tt(xtt).detach())))

    from videotuna.models.lvdm.modules.utils import GroupNormSpecific
    gn = GroupNormSpecific(32, 128)
    with torch.no_grad():
        gn.weight.copy_(RN(128, generator=gen))
        gn.bias.copy_(RN(128, generator=gen))
    round_params(gn)
    xg = r16(RN(2, 128, 6, 10, generator=gen) * 2 + 0.5)
    save("lvdm_groupnorm", dict(weight=gn.weight.detach().clone(), bias=gn.bias.detach().clone(), x=xg, eps=gn.eps,
                                out=gn(xg).detach(), out_silu=torch.nn.functional.silu(gn(xg)).detach()))


def golden_lvdm_temporal():
    """Temporal self-attention as VideoCrafter2/DynamiCrafter configure it (no relative position): plain and with the
    causal mask TemporalTransformer builds (attention.py:487-489). Own generator so earlier fixtures stay unchanged."""
    from videotuna.models.lvdm.modules import attention as A
    gen = _seeded(7)
    cases = {}
    kw = dict(query_dim=128, context_dim=None, heads=2, dim_head=64, temporal_length=16)
    m = round_params(A.CrossAttention(**kw))
    xt = RN(40, 16, 128, generator=gen)
    mask = torch.tril(torch.ones(1, 16, 16))
    cases["temporal"] = dict(kw=kw, sd=sd(m), x=xt, context=None, mask=None, out=m(xt).detach())
    cases["temporal_causal"] = dict(kw=kw, sd=sd(m), x=xt, context=None, mask=mask,
                                    out=m(xt, mask=mask.expand(40, -1, -1)).detach())
    xg = xt.clone().requires_grad_(True)
    do = RN(40, 16, 128, generator=gen)
    with torch.enable_grad():
        out = m(xg, mask=mask.expand(40, -1, -1))
        out.backward(do)
    cases["temporal_causal"].update(dout=do, dx=xg.grad.detach(),
                                    dw_q=m.to_q.weight.grad.detach().clone(), dw_v=m.to_v.weight.grad.detach().clone())
    save("lvdm_temporal_attention", cases)


def golden_hunyuan():
    gen = _seeded(1)
    M = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.models")
    att = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.attenion")
    pos = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.posemb_layers")
    nrm = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.norm_layers")
    mod = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.modulate_layers")

    B, S, H, D = 2, 150, 2, 128
    q, k, v = (RN(B, S, H, D, generator=gen) for _ in range(3))
    out_plain = att.attention(q, k, v, mode="torch")
    # two-segment block-diagonal mask = what mode="flash" computes through cu_seqlens (valid text 20 of 30, img 120)
    img_len, txt_len = 120, 30
    text_mask = torch.zeros(B, txt_len, dtype=torch.long)
    text_mask[0, :20] = 1
    text_mask[1, :30] = 1
    seg = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):
        seg[b, img_len + int(text_mask[b].sum()):] = 1
    mask = (seg[:, :, None] == seg[:, None, :])[:, None]
    out_mask = att.attention(q, k, v, mode="torch", attn_mask=mask)
    save("hunyuan_attention", dict(q=q, k=k, v=v, out_plain=out_plain, text_mask=text_mask, img_len=img_len,
                                   attn_mask=mask, out_mask=out_mask))

    rn = nrm.RMSNorm(D, elementwise_affine=True, eps=1e-6)
    with torch.no_grad():
        rn.weight.copy_(1 + 0.1 * RN(D, generator=gen))
    round_params(rn)
    cos, sin = pos.get_nd_rotary_pos_embed([16, 56, 56], (3, 4, 10), theta=256, use_real=True, theta_rescale_factor=1)
    xq, xk = RN(1, 120, H, D, generator=gen), RN(1, 120, H, D, generator=gen)
    rq, rk = pos.apply_rotary_emb(rn(xq), rn(xk), (cos, sin), head_first=False)
    shift, scale, gate = (RN(B, 64, generator=gen) for _ in range(3))
    xm = RN(B, 10, 64, generator=gen)
    save("hunyuan_norm_rope", dict(w=rn.weight.detach().clone(), eps=1e-6, cos=cos, sin=sin, xq=xq, xk=xk,
                                   norm_q=rn(xq).detach(), rope_q=rq.detach(), rope_k=rk.detach(), xm=xm, shift=shift,
                                   scale=scale, gate=gate, modulated=mod.modulate(xm, shift=shift, scale=scale),
                                   gated=mod.apply_gate(xm, gate=gate), rope_sizes=(3, 4, 10),
                                   rope_dim_list=[16, 56, 56], theta=256))

    hidden, heads = 256, 2
    dbl = M.MMDoubleStreamBlock(hidden, heads, mlp_width_ratio=1.0, qk_norm=True, qk_norm_type="rms", qkv_bias=True)
    sgl = M.MMSingleStreamBlock(hidden, heads, mlp_width_ratio=1.0, qk_norm=True, qk_norm_type="rms")
    dezero(dbl, gen)
    dezero(sgl, gen)
    with torch.no_grad():
        for mm in (dbl, sgl):
            for n_, p_ in mm.named_parameters():
                if n_.endswith("norm.weight"):
                    p_.copy_(1 + 0.1 * RN(p_.shape, generator=gen))
    round_params(dbl)
    round_params(sgl)
    # patch the name the blocks call so CPU runs take mode="torch" with the same segment mask flash would use
    img = RN(1, 120, hidden, generator=gen)
    txt = RN(1, 30, hidden, generator=gen)
    vec = RN(1, hidden, generator=gen)
    cu = torch.tensor([0, 140, 150], dtype=torch.int32)
    segm = torch.zeros(150, dtype=torch.long)
    segm[140:] = 1
    bmask = (segm[:, None] == segm[None, :])[None, None]
    orig = M.attention
    M.attention = lambda q_, k_, v_, **kw: orig(q_, k_, v_, mode="torch", attn_mask=bmask)
    try:
        img_o, txt_o = dbl(img, txt, vec, cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=150, max_seqlen_kv=150,
                           freqs_cis=(cos, sin))
        x_cat = torch.cat([img, txt], 1)
        sgl_o = sgl(x_cat, vec, 30, cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=150, max_seqlen_kv=150,
                    freqs_cis=(cos, sin))
    finally:
        M.attention = orig
    save("hunyuan_blocks", dict(hidden=hidden, heads=heads, mlp_width_ratio=1.0, dbl_sd=sd(dbl), sgl_sd=sd(sgl), img=img,
                                txt=txt, vec=vec, cu_seqlens=cu, cos=cos, sin=sin, img_out=img_o.detach(),
                                txt_out=txt_o.detach(), single_out=sgl_o.detach()))


def golden_wan():
    gen = _seeded(2)
    pkg = types.ModuleType("wan_ref_modules")
    pkg.__path__ = [os.path.join(REF, "videotuna/models/wan/wan/modules")]
    sys.modules["wan_ref_modules"] = pkg
    attn = importlib.import_module("wan_ref_modules.attention")
    model = importlib.import_module("wan_ref_modules.model")

    B, L, N, D = 2, 60, 2, 128
    q, k, v = (RN(B, L, N, D, generator=gen) for _ in range(3))
    # flash_attn is importable in this image but needs CUDA; clear the reference's own availability flags so that
    # attention() takes its SDPA fallback branch (attention.py:164-179), as it would on a machine without flash-attn.
    attn.FLASH_ATTN_2_AVAILABLE = False
    attn.FLASH_ATTN_3_AVAILABLE = False
    out = attn.attention(q, k, v)
    freqs = torch.cat([model.rope_params(1024, D - 4 * (D // 6)), model.rope_params(1024, 2 * (D // 6)),
                       model.rope_params(1024, 2 * (D // 6))], dim=1)
    grid = torch.tensor([[3, 4, 5], [2, 4, 5]])
    roped = model.rope_apply(q, grid, freqs)
    rms = model.WanRMSNorm(N * D, eps=1e-6)
    with torch.no_grad():
        rms.weight.copy_(1 + 0.1 * RN(N * D, generator=gen))
    round_params(rms)
    xr = RN(B, L, N * D, generator=gen)
    ln = model.WanLayerNorm(N * D, eps=1e-6)
    save("wan_ops", dict(q=q, k=k, v=v, sdpa_out=out,
                         grid=grid, roped=roped, rms_w=rms.weight.detach().clone(), xr=xr, rms_out=rms(xr).detach(),
                         ln_out=ln(xr).detach()))

    # one WanAttentionBlock; flash_attention rebound to the reference's own SDPA fallback body (needs no CUDA)
    def sdpa_flash(q, k, v, q_lens=None, k_lens=None, dropout_p=0.0, softmax_scale=None, q_scale=None, causal=False,
                   window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
        o = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2).contiguous()

    model.flash_attention = sdpa_flash
    dim, heads, ffn = 256, 2, 512
    blk = model.WanAttentionBlock("t2v_cross_attn", dim, ffn, heads, window_size=(-1, -1), qk_norm=True,
                                  cross_attn_norm=True, eps=1e-6)
    dezero(blk, gen)
    with torch.no_grad():
        blk.modulation.copy_(RN(1, 6, dim, generator=gen) / dim ** 0.5)
    round_params(blk)
    Lb = 3 * 4 * 5
    x = RN(1, Lb, dim, generator=gen)
    e = r16(RN(1, 6, dim, generator=gen) * 0.1)
    ctx = RN(1, 20, dim, generator=gen)
    freqs_b = torch.cat([model.rope_params(1024, 128 - 4 * (128 // 6)), model.rope_params(1024, 2 * (128 // 6)),
                         model.rope_params(1024, 2 * (128 // 6))], dim=1)
    y = blk(x, e, torch.tensor([Lb]), torch.tensor([[3, 4, 5]]), freqs_b, ctx, None)
    save("wan_block", dict(dim=dim, heads=heads, ffn=ffn, sd=sd(blk), x=x, e=e, context=ctx, grid=torch.tensor([[3, 4, 5]]),
                           out=y.detach()))


def golden_block_grads():
    """Forward + backward of whole reference blocks (fp32, CPU) on the state dicts and inputs of the fixtures above, plus
    a ResBlock: outputs, input gradients and a few parameter gradients for the block-level drop-ins (b200vt.blocks).
    Own generator; reads hunyuan_blocks.pt / wan_block.pt / lvdm_transformers.pt, so run after the others."""
    gen = _seeded(11)
    out = {}

    def load(name):
        return torch.load(os.path.join(OUT, name + ".pt"))

    def f32(t):
        return t.float() if t.is_floating_point() else t

    def grads_of(mod, names):
        # large weight gradients are stored in bf16 (they are compared by cosine similarity); small ones stay fp32
        params = dict(mod.named_parameters())
        return {n: (params[n].grad.detach().to(torch.bfloat16) if params[n].numel() > 16384 else params[n].grad.detach().clone())
                for n in names}

    with torch.enable_grad():
        # ---- Hunyuan double / single stream blocks -----------------------------------------------------------
        M = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.models")
        g = load("hunyuan_blocks")
        dbl = M.MMDoubleStreamBlock(g["hidden"], g["heads"], mlp_width_ratio=g["mlp_width_ratio"], qk_norm=True,
                                    qk_norm_type="rms", qkv_bias=True)
        dbl.load_state_dict({k: f32(v) for k, v in g["dbl_sd"].items()})
        sgl = M.MMSingleStreamBlock(g["hidden"], g["heads"], mlp_width_ratio=g["mlp_width_ratio"], qk_norm=True,
                                    qk_norm_type="rms")
        sgl.load_state_dict({k: f32(v) for k, v in g["sgl_sd"].items()})
        cu = g["cu_seqlens"]
        segm = torch.zeros(150, dtype=torch.long)
        segm[int(cu[1]):] = 1
        bmask = (segm[:, None] == segm[None, :])[None, None]
        orig = M.attention
        M.attention = lambda q_, k_, v_, **kw: orig(q_, k_, v_, mode="torch", attn_mask=bmask)
        try:
            img, txt, vec = (f32(g[k]).clone().requires_grad_(True) for k in ("img", "txt", "vec"))
            fc = (f32(g["cos"]), f32(g["sin"]))
            io, to = dbl(img, txt, vec, cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=150, max_seqlen_kv=150, freqs_cis=fc)
            d_io, d_to = RN(*io.shape, generator=gen), RN(*to.shape, generator=gen)
            torch.autograd.backward([io, to], [d_io, d_to])
            out["hunyuan_double"] = dict(d_img_out=d_io, d_txt_out=d_to, d_img=img.grad.clone(), d_txt=txt.grad.clone(),
                                         d_vec=vec.grad.clone(),
                                         **grads_of(dbl, ["img_attn_qkv.weight", "img_attn_q_norm.weight",
                                                          "txt_attn_k_norm.weight", "img_mod.linear.weight",
                                                          "txt_attn_proj.bias"]))
            x = torch.cat([f32(g["img"]), f32(g["txt"])], 1).clone().requires_grad_(True)
            vec2 = f32(g["vec"]).clone().requires_grad_(True)
            so = sgl(x, vec2, 30, cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=150, max_seqlen_kv=150, freqs_cis=fc)
            d_so = RN(*so.shape, generator=gen)
            so.backward(d_so)
            out["hunyuan_single"] = dict(d_out=d_so, d_x=x.grad.clone(), d_vec=vec2.grad.clone(),
                                         **grads_of(sgl, ["linear1.weight", "q_norm.weight", "k_norm.weight",
                                                          "modulation.linear.bias"]))
        finally:
            M.attention = orig

        # ---- Wan attention block -----------------------------------------------------------------------------
        pkg = types.ModuleType("wan_ref_modules")
        pkg.__path__ = [os.path.join(REF, "videotuna/models/wan/wan/modules")]
        sys.modules["wan_ref_modules"] = pkg
        model = importlib.import_module("wan_ref_modules.model")

        def sdpa_flash(q, k, v, q_lens=None, k_lens=None, dropout_p=0.0, softmax_scale=None, q_scale=None, causal=False,
                       window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
            o = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
            return o.transpose(1, 2).contiguous()

        model.flash_attention = sdpa_flash
        w = load("wan_block")
        blk = model.WanAttentionBlock("t2v_cross_attn", w["dim"], w["ffn"], w["heads"], window_size=(-1, -1), qk_norm=True,
                                      cross_attn_norm=True, eps=1e-6)
        blk.load_state_dict({k: f32(v) for k, v in w["sd"].items()})
        x = f32(w["x"]).clone().requires_grad_(True)
        e = f32(w["e"]).clone().requires_grad_(True)
        ctx = f32(w["context"]).clone().requires_grad_(True)
        freqs_b = torch.cat([model.rope_params(1024, 128 - 4 * (128 // 6)), model.rope_params(1024, 2 * (128 // 6)),
                             model.rope_params(1024, 2 * (128 // 6))], dim=1)
        y = blk(x, e, torch.tensor([x.shape[1]]), w["grid"], freqs_b, ctx, None)
        assert torch.allclose(y, f32(w["out"]), atol=1e-5)
        d_y = RN(*y.shape, generator=gen)
        y.backward(d_y)
        out["wan_block"] = dict(d_out=d_y, d_x=x.grad.clone(), d_e=e.grad.clone(), d_context=ctx.grad.clone(),
                                **grads_of(blk, ["modulation", "self_attn.norm_q.weight", "self_attn.q.weight",
                                                 "cross_attn.norm_k.weight", "norm3.weight", "ffn.0.bias"]))

        # ---- lvdm Spatial / Temporal transformer, ResBlock ---------------------------------------------------
        from videotuna.models.lvdm.modules import attention as A
        from videotuna.models.lvdm.modules.networks import openaimodel3d as O3
        lv = load("lvdm_transformers")
        st = A.SpatialTransformer(**lv["spatial"]["kw"])
        st.load_state_dict({k: f32(v) for k, v in lv["spatial"]["sd"].items()})
        xs = f32(lv["spatial"]["x"]).clone().requires_grad_(True)
        cs = f32(lv["spatial"]["context"]).clone().requires_grad_(True)
        ys = st(xs, cs)
        d_ys = RN(*ys.shape, generator=gen)
        ys.backward(d_ys)
        out["lvdm_spatial"] = dict(d_out=d_ys, d_x=xs.grad.clone(), d_context=cs.grad.clone(),
                                   **grads_of(st, ["norm.weight", "proj_in.weight", "transformer_blocks.0.norm2.weight",
                                                   "transformer_blocks.0.attn2.to_k.weight"]))
        tt = A.TemporalTransformer(**lv["temporal"]["kw"])
        tt.load_state_dict({k: f32(v) for k, v in lv["temporal"]["sd"].items()})
        xt = f32(lv["temporal"]["x"]).clone().requires_grad_(True)
        yt = tt(xt)
        d_yt = RN(*yt.shape, generator=gen)
        yt.backward(d_yt)
        out["lvdm_temporal"] = dict(d_out=d_yt, d_x=xt.grad.clone(),
                                    **grads_of(tt, ["norm.bias", "transformer_blocks.0.norm1.weight",
                                                    "transformer_blocks.0.attn1.to_q.weight"]))
        rkw = dict(channels=64, emb_channels=96, dropout=0.0, out_channels=128, dims=2, use_checkpoint=False,
                   use_temporal_conv=False)
        rb = O3.ResBlock(**rkw)
        dezero(rb, gen)
        with torch.no_grad():
            for n_, p_ in rb.named_parameters():
                if n_ in ("in_layers.0.weight", "out_layers.0.weight"):
                    p_.copy_(1 + 0.1 * RN(p_.shape, generator=gen))
                if n_ in ("in_layers.0.bias", "out_layers.0.bias"):
                    p_.copy_(0.1 * RN(p_.shape, generator=gen))
        round_params(rb)
        xr = r16(RN(2, 64, 6, 10, generator=gen) * 1.5 + 0.3).requires_grad_(True)
        er = RN(2, 96, generator=gen).requires_grad_(True)
        yr = rb(xr, er)
        d_yr = RN(*yr.shape, generator=gen)
        yr.backward(d_yr)
        out["lvdm_resblock"] = dict(kw=rkw, sd=sd(rb), x=xr.detach().clone(), emb=er.detach().clone(), out=yr.detach(),
                                    d_out=d_yr, d_x=xr.grad.clone(), d_emb=er.grad.clone(),
                                    **grads_of(rb, ["in_layers.0.weight", "out_layers.0.bias", "in_layers.2.weight"]))
    save("block_grads", out)


def golden_hunyuan_i2v():
    """The i2v twins of the Hunyuan blocks (hyvideo_i2v/modules/models.py:136-297, 371-462) called the way
    HYVideoDiffusionTransformer.forward calls them — 11 positional arguments (:749-761, 776-788) — once as T2V
    (condition_type None) and once with the "token_replace" first-frame modulation (modulate_layers.py:37-96), forward and
    backward."""
    gen = _seeded(21)
    M = importlib.import_module("videotuna.models.hunyuan.hyvideo_i2v.modules.models")
    pos = importlib.import_module("videotuna.models.hunyuan.hyvideo_i2v.modules.posemb_layers")
    hidden, heads, ff = 128, 2, 40  # head dim 64; 40 = one latent frame of the (3, 4, 10) token grid
    dbl = M.MMDoubleStreamBlock(hidden, heads, mlp_width_ratio=1.0, qk_norm=True, qk_norm_type="rms", qkv_bias=True)
    sgl = M.MMSingleStreamBlock(hidden, heads, mlp_width_ratio=1.0, qk_norm=True, qk_norm_type="rms")
    dezero(dbl, gen)
    dezero(sgl, gen)
    with torch.no_grad():
        for mm in (dbl, sgl):
            for n_, p_ in mm.named_parameters():
                if n_.endswith("norm.weight"):
                    p_.copy_(1 + 0.1 * RN(p_.shape, generator=gen))
    round_params(dbl)
    round_params(sgl)
    cos, sin = pos.get_nd_rotary_pos_embed([8, 28, 28], (3, 4, 10), theta=256, use_real=True, theta_rescale_factor=1)
    cu = torch.tensor([0, 140, 150], dtype=torch.int32)
    segm = torch.zeros(150, dtype=torch.long)
    segm[140:] = 1
    bmask = (segm[:, None] == segm[None, :])[None, None]
    orig = M.attention
    M.attention = lambda q_, k_, v_, **kw: orig(q_, k_, v_, mode="torch", attn_mask=bmask)
    out = dict(hidden=hidden, heads=heads, mlp_width_ratio=1.0, dbl_sd=sd(dbl), sgl_sd=sd(sgl), cu_seqlens=cu, cos=cos,
               sin=sin, first_frame_tokens=ff)
    pnames_d = ["img_attn_qkv.weight", "img_mod.linear.bias", "img_attn_k_norm.weight"]
    pnames_s = ["linear1.bias", "modulation.linear.bias", "linear2.weight"]
    try:
        with torch.enable_grad():
            base = dict(img=RN(1, 120, hidden, generator=gen), txt=RN(1, 30, hidden, generator=gen),
                        vec=RN(1, hidden, generator=gen), trv=RN(1, hidden, generator=gen))
            d_io, d_to = RN(1, 120, hidden, generator=gen), RN(1, 30, hidden, generator=gen)
            d_so = RN(1, 150, hidden, generator=gen)
            out.update(base, d_img_out=d_io, d_txt_out=d_to, d_single_out=d_so)
            for tag, cond in (("t2v", None), ("tr", "token_replace")):
                img, txt, vec, trv = (base[k].clone().requires_grad_(True) for k in ("img", "txt", "vec", "trv"))
                dbl.zero_grad()
                io, to = dbl(img, txt, vec, cu, cu, 150, 150, (cos, sin), cond, trv if cond else None, ff if cond else None)
                torch.autograd.backward([io, to], [d_io, d_to])
                rec = dict(img_out=io.detach(), txt_out=to.detach(), d_img=img.grad.clone(), d_txt=txt.grad.clone(),
                           d_vec=vec.grad.clone(),
                           **{n: dict(dbl.named_parameters())[n].grad.detach().to(torch.bfloat16) for n in pnames_d})
                if cond:
                    rec["d_trv"] = trv.grad.clone()
                out["double_" + tag] = rec
                x = torch.cat([base["img"], base["txt"]], 1).clone().requires_grad_(True)
                vec2, trv2 = base["vec"].clone().requires_grad_(True), base["trv"].clone().requires_grad_(True)
                sgl.zero_grad()
                so = sgl(x, vec2, 30, cu, cu, 150, 150, (cos, sin), cond, trv2 if cond else None, ff if cond else None)
                so.backward(d_so)
                rec = dict(out=so.detach(), d_x=x.grad.clone(), d_vec=vec2.grad.clone(),
                           **{n: dict(sgl.named_parameters())[n].grad.detach().to(torch.bfloat16) for n in pnames_s})
                if cond:
                    rec["d_trv"] = trv2.grad.clone()
                out["single_" + tag] = rec
    finally:
        M.attention = orig
    save("hunyuan_i2v_blocks", out)


def golden_wan_i2v_cross():
    """WanI2VCrossAttention (wan/modules/model.py:184-225): 257 CLIP image tokens + text tokens, forward and backward."""
    gen = _seeded(22)
    pkg = types.ModuleType("wan_ref_modules")
    pkg.__path__ = [os.path.join(REF, "videotuna/models/wan/wan/modules")]
    sys.modules["wan_ref_modules"] = pkg
    model = importlib.import_module("wan_ref_modules.model")

    def sdpa_flash(q, k, v, q_lens=None, k_lens=None, dropout_p=0.0, softmax_scale=None, q_scale=None, causal=False,
                   window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
        mask = None
        if k_lens is not None:  # flash_attention packs the first k_lens[b] keys of every sample (attention.py:62-71)
            mask = (torch.arange(k.shape[1])[None, :] < k_lens[:, None])[:, None, None, :]
        o = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                             attn_mask=mask)
        return o.transpose(1, 2).contiguous()

    model.flash_attention = sdpa_flash
    dim, heads = 256, 2
    m = model.WanI2VCrossAttention(dim, heads, qk_norm=True, eps=1e-6)
    with torch.no_grad():
        for n_, p_ in m.named_parameters():
            if n_.startswith("norm") and n_.endswith("weight"):
                p_.copy_(1 + 0.1 * RN(p_.shape, generator=gen))
    round_params(m)
    x = RN(2, 60, dim, generator=gen).requires_grad_(True)
    ctx = RN(2, 257 + 24, dim, generator=gen).requires_grad_(True)
    lens = torch.tensor([24, 17])
    d_y = RN(2, 60, dim, generator=gen)
    with torch.enable_grad():
        y = m(x, ctx, lens)
        y.backward(d_y)
    params = dict(m.named_parameters())
    save("wan_i2v_cross", dict(dim=dim, heads=heads, sd=sd(m), x=x.detach().clone(), context=ctx.detach().clone(),
                               context_lens=lens, out=y.detach(), d_out=d_y, d_x=x.grad.clone(),
                               d_context=ctx.grad.clone(),
                               **{n: params[n].grad.detach().clone() for n in ("norm_k_img.weight", "norm_q.weight")},
                               **{n: params[n].grad.detach().to(torch.bfloat16) for n in ("k_img.weight", "v.weight")}))


def golden_lvdm_extra():
    """TemporalConvBlock (openaimodel3d.py:258-310: four GroupNorm(32) + SiLU + Conv3d stages) forward + backward, and the
    VideoCrafter1 relative-position temporal attention (attention.py:19-42, 129-133, 145-148) forward + backward."""
    gen = _seeded(23)
    from videotuna.models.lvdm.modules import attention as A
    from videotuna.models.lvdm.modules.networks import openaimodel3d as O3
    tc = O3.TemporalConvBlock(64, out_channels=64, dropout=0.0)
    dezero(tc, gen)
    with torch.no_grad():
        for n_, p_ in tc.named_parameters():
            if n_.endswith(".0.weight"):
                p_.copy_(1 + 0.1 * RN(p_.shape, generator=gen))
            if n_.endswith(".0.bias"):
                p_.copy_(0.1 * RN(p_.shape, generator=gen))
    round_params(tc)
    x = r16(RN(2, 64, 8, 5, 6, generator=gen) * 1.5 + 0.3).requires_grad_(True)
    d_y = RN(2, 64, 8, 5, 6, generator=gen)
    with torch.enable_grad():
        y = tc(x)
        y.backward(d_y)
    params = dict(tc.named_parameters())
    out = dict(tconv=dict(channels=64, sd=sd(tc), x=x.detach().clone(), out=y.detach(), d_out=d_y, d_x=x.grad.clone(),
                          **{n: params[n].grad.detach().clone() for n in ("conv1.0.weight", "conv3.0.bias", "conv2.3.bias")}))

    kw = dict(query_dim=128, context_dim=None, heads=2, dim_head=64, relative_position=True, temporal_length=16)
    m = A.CrossAttention(**kw)
    round_params(m)
    xt = RN(12, 16, 128, generator=gen)
    mask = torch.tril(torch.ones(1, 16, 16))
    d_o = RN(12, 16, 128, generator=gen)
    rel = {}
    with torch.enable_grad():
        for tag, mk in (("plain", None), ("causal", mask.expand(12, -1, -1))):
            xg = xt.clone().requires_grad_(True)
            m.zero_grad()
            o = m(xg, mask=mk)
            o.backward(d_o)
            p_ = dict(m.named_parameters())
            rel[tag] = dict(out=o.detach(), d_x=xg.grad.clone(),
                            d_rel_k=p_["relative_position_k.embeddings_table"].grad.detach().clone(),
                            d_rel_v=p_["relative_position_v.embeddings_table"].grad.detach().clone(),
                            d_to_q=p_["to_q.weight"].grad.detach().clone())
    out["relpos"] = dict(kw=kw, sd=sd(m), x=xt, mask=mask, d_out=d_o, **rel)
    save("lvdm_extra", out)


def main():
    global CHECK
    CHECK = "--check" in sys.argv
    install_shims()
    torch.set_grad_enabled(False)
    golden_lvdm()
    golden_hunyuan()
    golden_wan()
    golden_lvdm_temporal()
    golden_block_grads()
    golden_hunyuan_i2v()
    golden_wan_i2v_cross()
    golden_lvdm_extra()
    if CHECK:
        print("check:", "all fixtures regenerate bit-identically" if not MISMATCHES else f"{len(MISMATCHES)} mismatches")
        raise SystemExit(1 if MISMATCHES else 0)


if __name__ == "__main__":
    main()
