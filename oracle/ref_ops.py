"""oracle/ref_ops.py — CPU restatement of the reference's attention hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it. The product path (videotuna-dev_b200/) never does, and has no CPU fallback.

Every function restates one reference callable in plain PyTorch on whatever dtype it is given (tests use fp32/fp64 on
CPU) and cites the reference lines it follows (paths relative to /root/reference/videotuna/models/).

Pinning: the reference ships no tests or golden vectors for this path, so the oracle is pinned
against outputs of the reference's own code imported in the development container:
tests/golden/make_golden.py runs the reference modules on seeded inputs and commits inputs+outputs as
tests/golden/*.pt; tests/test_oracle_golden.py checks every function below against them.
Third-party arithmetic the reference calls but does not vendor (diffusers 0.32.2 CogVideoX/Hunyuan processors,
xfuser 0.4.3.post2, flash-attn) is "parity unpinned": see DESIGN.md.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


# ---------------------------------------------------------------------------------------------------------------------
# attention cores
# ---------------------------------------------------------------------------------------------------------------------
def lvdm_attention_core(q: Tensor, k: Tensor, v: Tensor, scale: float, rel_k: Optional[Tensor] = None,
                        rel_v: Optional[Tensor] = None, mask: Optional[Tensor] = None) -> Tensor:
    """lvdm/modules/attention.py:128-148.  q (BH,N,D), k/v (BH,M,D); rel_k/rel_v (N,M,D); mask (BH|1,N,M) >0.5 = keep.

    sim = einsum(q,k)*scale [+ einsum(q,k2)*scale]; masked_fill(~mask, -finfo.max); softmax; out = sim@v [+ sim@v2].
    """
    sim = torch.einsum("bid,bjd->bij", q, k) * scale
    if rel_k is not None:
        sim = sim + torch.einsum("btd,tsd->bts", q, rel_k) * scale
    if mask is not None:
        sim = sim.masked_fill(~(mask > 0.5), -torch.finfo(sim.dtype).max)
    sim = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", sim, v)
    if rel_v is not None:
        out = out + torch.einsum("bts,tsd->btd", sim, rel_v)
    return out


def lvdm_split_heads(t: Tensor, h: int) -> Tensor:
    """rearrange 'b n (h d) -> (b h) n d'  (attention.py:126)."""
    b, n, hd = t.shape
    return t.view(b, n, h, hd // h).permute(0, 2, 1, 3).reshape(b * h, n, hd // h)


def lvdm_merge_heads(t: Tensor, h: int) -> Tensor:
    """rearrange '(b h) n d -> b n (h d)'  (attention.py:149)."""
    bh, n, d = t.shape
    return t.view(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)


def lvdm_relative_position(table: Tensor, length_q: int, length_k: int, max_rel: int) -> Tensor:
    """RelativePosition.forward, attention.py:31-42: table[(clamp(k - q, -max, max) + max)] -> (Nq, Nk, D)."""
    rq = torch.arange(length_q, device=table.device)
    rk = torch.arange(length_k, device=table.device)
    dist = (rk[None, :] - rq[:, None]).clamp(-max_rel, max_rel) + max_rel
    return table[dist.long()]


def sdpa_blhd(q: Tensor, k: Tensor, v: Tensor, attn_mask: Optional[Tensor] = None,
              scale: Optional[float] = None) -> Tensor:
    """softmax(q k^T * scale [+mask]) v on (B,L,H,D) tensors, returning (B,Lq,H,D). Explicit (no fused kernel) so it
    runs identically in fp32 and fp64."""
    scale = 1.0 / math.sqrt(q.shape[-1]) if scale is None else scale
    s = torch.einsum("bihd,bjhd->bhij", q, k) * scale
    if attn_mask is not None:
        if attn_mask.dtype == torch.bool:
            s = s.masked_fill(~attn_mask, float("-inf"))
        else:
            s = s + attn_mask
    p = s.softmax(dim=-1)
    return torch.einsum("bhij,bjhd->bihd", p, v)


def attention_fwd_bwd_chunked(q: Tensor, k: Tensor, v: Tensor, do: Tensor, scale: Optional[float] = None,
                              segments: Optional[Sequence[int]] = None, k_len: Optional[int] = None,
                              chunk: int = 2048) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """sdpa_blhd + its analytic backward for ONE (sample, head): q, do (Lq, D), k, v (Lk, D) fp32 -> (o, lse, dq, dk, dv),
    evaluated in chunks of `chunk` query rows so that the (Lq, Lk) score matrix never exists at once. This is the same
    arithmetic as the reference's torch path (lvdm attention.py:128-144, hunyuan attenion.py:101-106 + autograd:
    P = softmax(S), O = P V, dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(dO o O)), dQ = dS K s, dK = dS^T Q s); it
    exists so that the BASELINE.json sizes (119 056 tokens: a 57 GB score matrix per head) can be compared on the GPU
    box in fp32 (tests/test_gpu_attention_fullsize.py). Runs on whatever device the inputs live on; pinned to
    sdpa_blhd + autograd on CPU by tests/test_oracle_golden.py.
    segments: packed varlen boundaries (cu_seqlens of attenion.py:34-57) — row i attends key j iff same segment, rows /
    keys outside every segment give zeros; k_len: keys >= k_len are masked (wan attention.py:62-71)."""
    Lq, D = q.shape
    Lk = k.shape[0]
    scale = 1.0 / math.sqrt(D) if scale is None else scale
    dev = q.device
    seg_q = seg_k = None
    if segments is not None:
        cu = [int(c) for c in segments]
        seg_q = torch.full((Lq,), -1, dtype=torch.long, device=dev)
        seg_k = torch.full((Lk,), -2, dtype=torch.long, device=dev)
        for s_ in range(len(cu) - 1):
            seg_q[cu[s_]:cu[s_ + 1]] = s_
            seg_k[cu[s_]:cu[s_ + 1]] = s_
    o = torch.zeros_like(q)
    lse = torch.full((Lq,), float("-inf"), dtype=q.dtype, device=dev)
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    for r0 in range(0, Lq, chunk):
        r1 = min(Lq, r0 + chunk)
        s = (q[r0:r1] @ k.T) * scale
        if seg_q is not None:
            s = s.masked_fill(seg_q[r0:r1, None] != seg_k[None, :], float("-inf"))
        if k_len is not None:
            s[:, k_len:] = float("-inf")
        l = torch.logsumexp(s, dim=-1)
        p = torch.exp(s - l.masked_fill(torch.isinf(l), 0.0)[:, None])  # fully masked rows: exp(-inf) = 0
        lse[r0:r1] = l
        oc = p @ v
        o[r0:r1] = oc
        doc = do[r0:r1]
        dv += p.T @ doc
        dp = doc @ v.T
        ds = p * (dp - (doc * oc).sum(dim=-1, keepdim=True))
        dq[r0:r1] = (ds @ k) * scale
        dk += (ds.T @ q[r0:r1]) * scale
    return o, lse, dq, dk, dv


def hunyuan_attention_torch(q: Tensor, k: Tensor, v: Tensor, attn_mask: Optional[Tensor] = None) -> Tensor:
    """hunyuan/hyvideo_t2v/modules/attenion.py:101-106,151-156 (mode="torch"): q,k,v (B,S,H,D) -> (B,S,H*D).
    pre: transpose(1,2); SDPA(attn_mask); post: transpose back; reshape(b, s, -1)."""
    b, s, h, d = q.shape
    return sdpa_blhd(q, k, v, attn_mask).reshape(b, s, h * d)


def hunyuan_attention_torch_fused(q: Tensor, k: Tensor, v: Tensor, attn_mask: Optional[Tensor] = None) -> Tensor:
    """The same mode="torch" branch (attenion.py:97-106,151-156) written with the call the reference itself makes,
    F.scaled_dot_product_attention, instead of the explicit softmax above. It never materialises the (S,S) score
    matrix, so it is the form bench.py times on the host cores at the full 119 056-token sequence
    (cpu_baseline / --impl reference). Checked against hunyuan_attention_torch in tests/test_oracle_golden.py."""
    b, s, h, d = q.shape
    if attn_mask is not None and attn_mask.dtype != torch.bool:
        attn_mask = attn_mask.to(q.dtype)
    x = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=attn_mask,
                                       dropout_p=0.0, is_causal=False)
    return x.transpose(1, 2).reshape(b, s, h * d)


def hunyuan_cu_seqlens(text_mask: Tensor, img_len: int) -> Tensor:
    """get_cu_seqlens, attenion.py:34-57: [0, img+valid_0, img+max_0, (img+max)+img+valid_1, 2*(img+max), ...]."""
    batch_size = text_mask.shape[0]
    text_len = text_mask.sum(dim=1)
    max_len = text_mask.shape[1] + img_len
    cu = torch.zeros([2 * batch_size + 1], dtype=torch.int32)
    for i in range(batch_size):
        s = int(text_len[i]) + img_len
        cu[2 * i + 1] = i * max_len + s
        cu[2 * i + 2] = (i + 1) * max_len
    return cu


def varlen_block_mask(cu_seqlens: Sequence[int], total: int) -> Tensor:
    """Block-diagonal bool mask (total,total) equal to what flash_attn_varlen_func computes for packed segments
    (attenion.py:108-119): token i attends token j iff they lie in the same [cu[s], cu[s+1]) segment."""
    seg = torch.full((total,), -1, dtype=torch.long)
    cu = [int(c) for c in cu_seqlens]
    for s in range(len(cu) - 1):
        seg[cu[s]:cu[s + 1]] = s
    return (seg[:, None] == seg[None, :]) & (seg[:, None] >= 0)


def hunyuan_attention_flash_semantics(q: Tensor, k: Tensor, v: Tensor, cu_seqlens: Tensor) -> Tensor:
    """mode="flash" (attenion.py:107-120) expressed with the block-diagonal mask: tensors (B,S,H,D) are flattened to
    (B*S,H,D) and attended per segment. Rows outside every segment (none, by construction) would be zero."""
    b, s, h, d = q.shape
    qf, kf, vf = (t.reshape(1, b * s, h, d) for t in (q, k, v))
    mask = varlen_block_mask(cu_seqlens.tolist(), b * s)[None, None]
    return sdpa_blhd(qf, kf, vf, mask).reshape(b, s, h * d)


def wan_flash_attention(q: Tensor, k: Tensor, v: Tensor, q_lens: Optional[Tensor] = None,
                        k_lens: Optional[Tensor] = None, softmax_scale: Optional[float] = None,
                        q_scale: Optional[float] = None) -> Tensor:
    """wan/wan/modules/attention.py:24-130 semantics (FA varlen over per-sample lengths), restated densely:
    q (B,Lq,N,C), k/v (B,Lk,N,C). Keys >= k_lens[b] are excluded; query rows >= q_lens[b] produce zeros (the reference
    unflattens the packed output into (b, lq): rows past q_lens are only defined when q_lens is None). Output has q's
    dtype (:57,130)."""
    out_dtype = q.dtype
    b, lq, lk = q.shape[0], q.shape[1], k.shape[1]
    if q_scale is not None:
        q = q * q_scale
    mask = None
    if k_lens is not None:
        mask = (torch.arange(lk)[None, :] < k_lens[:, None].cpu())[:, None, None, :]  # (B,1,1,Lk)
    out = sdpa_blhd(q, k, v, mask, softmax_scale)
    if q_lens is not None:
        keep = (torch.arange(lq)[None, :] < q_lens[:, None].cpu())[:, :, None, None]
        out = out * keep
    return out.to(out_dtype)


def wan_attention_sdpa_fallback(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """attention.py:171-179: transpose(1,2) -> F.scaled_dot_product_attention -> transpose back; lengths are ignored."""
    out = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    return out.transpose(1, 2).contiguous()


# ---------------------------------------------------------------------------------------------------------------------
# norms / rope / modulate
# ---------------------------------------------------------------------------------------------------------------------
def hunyuan_rmsnorm(x: Tensor, weight: Optional[Tensor], eps: float = 1e-6) -> Tensor:
    """norm_layers.py:33-59: (x.float() * rsqrt(mean(x^2,-1) + eps)).type_as(x) * weight."""
    out = (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps)).type_as(x)
    return out * weight if weight is not None else out


def hunyuan_rotate_half(x: Tensor) -> Tensor:
    """posemb_layers.py:133-137."""
    xr, xi = x.float().reshape(*x.shape[:-1], -1, 2).unbind(-1)
    return torch.stack([-xi, xr], dim=-1).flatten(3)


def hunyuan_apply_rotary_emb(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """posemb_layers.py:164-171 (tuple branch, head_first=False): x (B,S,H,D), cos/sin (S,D) ->
    (x.float()*cos + rotate_half(x.float())*sin).type_as(x)."""
    c = cos.view(1, cos.shape[0], 1, cos.shape[1])
    s = sin.view(1, sin.shape[0], 1, sin.shape[1])
    return (x.float() * c + hunyuan_rotate_half(x.float()) * s).type_as(x)


def hunyuan_1d_rope(dim: int, pos: Tensor, theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    """get_1d_rotary_pos_embed, posemb_layers.py:261-310 (use_real=True): cos/sin (S, dim) with repeat_interleave(2)."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    freqs = torch.outer(pos.float(), freqs)
    return freqs.cos().repeat_interleave(2, dim=1), freqs.sin().repeat_interleave(2, dim=1)


def hunyuan_nd_rope(rope_dim_list: Sequence[int], sizes: Sequence[int], theta: float = 256.0) -> Tuple[Tensor, Tensor]:
    """get_nd_rotary_pos_embed, posemb_layers.py:191-258 for start=sizes (meshgrid over (T,H,W), 'ij'), no
    interpolation: per-axis 1-D tables concatenated on the feature dim -> cos/sin (T*H*W, sum(rope_dim_list))."""
    grids = torch.meshgrid(*[torch.arange(n, dtype=torch.float32) for n in sizes], indexing="ij")
    cs, sn = [], []
    for i, d in enumerate(rope_dim_list):
        c, s = hunyuan_1d_rope(d, grids[i].reshape(-1), theta)
        cs.append(c)
        sn.append(s)
    return torch.cat(cs, dim=1), torch.cat(sn, dim=1)


def wan_rope_params(max_seq_len: int, dim: int, theta: float = 10000.0) -> Tensor:
    """wan/wan/modules/model.py:29-36: complex128 table (max_seq_len, dim/2)."""
    freqs = torch.outer(torch.arange(max_seq_len),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def wan_freqs_table(d: int) -> Tensor:
    """WanModel.__init__, model.py:469-474: cat of three axis tables for head dim d."""
    return torch.cat([wan_rope_params(1024, d - 4 * (d // 6)), wan_rope_params(1024, 2 * (d // 6)),
                      wan_rope_params(1024, 2 * (d // 6))], dim=1)


def wan_rope_apply(x: Tensor, grid_sizes: Tensor, freqs: Tensor) -> Tensor:
    """model.py:40-67: x (B,L,N,D); per sample complex multiply in float64 on the first f*h*w tokens; returns float32."""
    n, c = x.size(2), x.size(3) // 2
    fs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    output = []
    for i, (f, h, w) in enumerate(grid_sizes.tolist()):
        seq_len = f * h * w
        x_i = torch.view_as_complex(x[i, :seq_len].to(torch.float64).reshape(seq_len, n, -1, 2))
        freqs_i = torch.cat([
            fs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
            fs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
            fs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq_len, 1, -1)
        x_i = torch.view_as_real(x_i * freqs_i).flatten(2)
        x_i = torch.cat([x_i, x[i, seq_len:]])
        output.append(x_i)
    return torch.stack(output).float()


def wan_rope_cos_sin(grid_size: Sequence[int], freqs: Tensor) -> Tuple[Tensor, Tensor]:
    """The (L, D) cos/sin tables (repeat_interleave(2) form) equivalent to wan_rope_apply's per-token complex
    multipliers for one (f,h,w) grid — the host-side table b200vt builds for its fused kernel."""
    f, h, w = (int(v) for v in grid_size)
    c = freqs.shape[1]
    fs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    fr = torch.cat([
        fs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        fs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
        fs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, -1)
    return (fr.real.float().repeat_interleave(2, dim=1).contiguous(),
            fr.imag.float().repeat_interleave(2, dim=1).contiguous())


def wan_rmsnorm(x: Tensor, weight: Tensor, eps: float = 1e-6) -> Tensor:
    """WanRMSNorm.forward, model.py:78-86: _norm(x.float()).type_as(x) * weight over the last dim (= full model dim)."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * weight


def layer_norm(x: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], eps: float) -> Tensor:
    """nn.LayerNorm as used by lvdm BasicTransformerBlock (attention.py:262-264, eps 1e-5, affine), Hunyuan
    (models.py:63-65, eps 1e-6, no affine) and WanLayerNorm (model.py:89-99, float32 compute)."""
    return F.layer_norm(x.float(), (x.shape[-1],), None if gamma is None else gamma.float(),
                        None if beta is None else beta.float(), eps).type_as(x)


def modulate(x: Tensor, shift: Optional[Tensor] = None, scale: Optional[Tensor] = None) -> Tensor:
    """hunyuan modulate_layers.py:31-49."""
    if scale is None and shift is None:
        return x
    if shift is None:
        return x * (1 + scale.unsqueeze(1))
    if scale is None:
        return x + shift.unsqueeze(1)
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def apply_gate(x: Tensor, gate: Optional[Tensor] = None, tanh: bool = False) -> Tensor:
    """hunyuan modulate_layers.py:52-68."""
    if gate is None:
        return x
    return x * gate.unsqueeze(1).tanh() if tanh else x * gate.unsqueeze(1)


def ln_modulate(x: Tensor, gamma, beta, scale, shift, eps: float) -> Tensor:
    """modulate(LayerNorm(x), shift, scale): hunyuan models.py:161-164; wan model.py:294-296 (norm1(x).float()*(1+e1)+e0)."""
    return modulate(layer_norm(x, gamma, beta, eps), shift, scale)


def gate_residual(x: Tensor, branch: Tensor, gate: Optional[Tensor]) -> Tensor:
    """x + apply_gate(branch, gate): hunyuan models.py:231; wan model.py:298 (x + y * e2)."""
    return x + apply_gate(branch, gate)


def groupnorm_silu(x: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], groups: int, eps: float,
                   silu: bool) -> Tensor:
    """GroupNormSpecific.forward (lvdm/modules/utils.py:192-194: super().forward(x.float()).type(x.dtype)) followed by
    nn.SiLU when used as ResBlock.in_layers/out_layers (openaimodel3d.py:156-160,184-186)."""
    y = F.group_norm(x.float(), groups, None if gamma is None else gamma.float(),
                     None if beta is None else beta.float(), eps).type(x.dtype)
    return F.silu(y) if silu else y


# ---------------------------------------------------------------------------------------------------------------------
# error metrics used by the parity tests (tensor-normalised error; gradient cosine)
# ---------------------------------------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------------------------------
# diffusers 0.32.2 attention processors — PARITY UNPINNED (the wheel is neither vendored by the reference nor installed
# here): restated from the published algorithm; anchored on the reference's call sites cogvideo_hf/cogvideo_pl.py:123,
# 862-868 and hyvideo_t2v/hunyuanvideo.py:209, 946-955 and on the in-tree SAT description of the same model
# (cogvideo_sat/dit_video_concat.py:263-427: text tokens first and un-rotated, interleaved rotate_half, per-head
# LayerNorm of q and k).
# ---------------------------------------------------------------------------------------------------------------------
def diffusers_cogvideox_attention(hidden: Tensor, encoder_hidden: Tensor, wq: Tensor, wk: Tensor, wv: Tensor, bq, bk, bv,
                                  wo: Tensor, bo, heads: int, ln_q=None, ln_k=None, rotary=None, eps: float = 1e-6):
    """CogVideoXAttnProcessor2_0: x = [text; video]; q,k,v = Linear(x); per-head LayerNorm(q), LayerNorm(k); RoPE on the
    video tokens; SDPA; to_out; split back into (video, text). ln_q/ln_k: (weight, bias) or None; rotary: (cos, sin)."""
    T = encoder_hidden.shape[1]
    x = torch.cat([encoder_hidden, hidden], dim=1)
    B, S, _ = x.shape
    q = F.linear(x, wq, bq).view(B, S, heads, -1)
    k = F.linear(x, wk, bk).view(B, S, heads, -1)
    v = F.linear(x, wv, bv).view(B, S, heads, -1)
    D = q.shape[-1]
    if ln_q is not None:
        q = F.layer_norm(q, (D,), ln_q[0], ln_q[1], eps)
    if ln_k is not None:
        k = F.layer_norm(k, (D,), ln_k[0], ln_k[1], eps)
    if rotary is not None:
        q = torch.cat([q[:, :T], hunyuan_apply_rotary_emb(q[:, T:], *rotary)], dim=1)
        k = torch.cat([k[:, :T], hunyuan_apply_rotary_emb(k[:, T:], *rotary)], dim=1)
    o = sdpa_blhd(q, k, v).reshape(B, S, heads * D)
    o = F.linear(o, wo, bo)
    return o[:, T:], o[:, :T]


def diffusers_cogvideox_norm_zero(hidden: Tensor, encoder: Tensor, temb: Tensor, lin_w: Tensor, lin_b, ln_w, ln_b, eps: float):
    """CogVideoXLayerNormZero.forward (diffusers 0.32.2 models/normalization.py; PARITY UNPINNED as above):
    shift, scale, gate, enc_shift, enc_scale, enc_gate = Linear(SiLU(temb)).chunk(6); both streams through the same
    LayerNorm, then * (1 + scale) + shift; gates returned as (B, 1, C)."""
    shift, scale, gate, e_shift, e_scale, e_gate = F.linear(F.silu(temb), lin_w, lin_b).chunk(6, dim=1)
    C = hidden.shape[-1]
    h = F.layer_norm(hidden, (C,), ln_w, ln_b, eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    e = F.layer_norm(encoder, (C,), ln_w, ln_b, eps) * (1 + e_scale)[:, None, :] + e_shift[:, None, :]
    return h, e, gate[:, None, :], e_gate[:, None, :]


def diffusers_cogvideox_block(hidden: Tensor, encoder: Tensor, temb: Tensor, p: dict, heads: int, rotary=None):
    """CogVideoXBlock.forward (diffusers 0.32.2 models/transformers/cogvideox_transformer_3d.py; PARITY UNPINNED):
    norm1 -> attention (diffusers_cogvideox_attention) -> gated residuals -> norm2 -> FeedForward(gelu-approximate) over
    [text; video] -> gated residuals. p: norm1/norm2 = (lin_w, lin_b, ln_w, ln_b, eps); attn = kwargs of
    diffusers_cogvideox_attention; ff = (w1, b1, w2, b2)."""
    T = encoder.shape[1]
    hn, en, g, eg = diffusers_cogvideox_norm_zero(hidden, encoder, temb, *p["norm1"])
    ah, ae = diffusers_cogvideox_attention(hn, en, heads=heads, rotary=rotary, **p["attn"])
    hidden = hidden + g * ah
    encoder = encoder + eg * ae
    hn, en, g, eg = diffusers_cogvideox_norm_zero(hidden, encoder, temb, *p["norm2"])
    w1, b1, w2, b2 = p["ff"]
    ff = F.linear(F.gelu(F.linear(torch.cat([en, hn], dim=1), w1, b1), approximate="tanh"), w2, b2)
    hidden = hidden + g * ff[:, T:]
    encoder = encoder + eg * ff[:, :T]
    return hidden, encoder


def diffusers_hunyuan_attention(hidden: Tensor, encoder_hidden: Tensor, p: dict, heads: int, rotary=None,
                                valid_len: Optional[Tensor] = None, eps: float = 1e-6):
    """HunyuanVideoAttnProcessor2_0, double-stream form (p has add_q/add_k/add_v projections) or single-stream form
    (hidden already is [video; text], encoder_hidden marks the text length). Per-head RMSNorm weights p["norm_q"] etc.;
    RoPE on the video tokens; SDPA under the reference's square validity mask (i < valid) & (j < valid)
    (hunyuanvideo.py builds it from the prompt mask); to_out / to_add_out when present."""
    double = "add_q" in p
    T = encoder_hidden.shape[1]
    x = hidden if double else torch.cat([hidden, encoder_hidden], dim=1)
    B, S1, _ = x.shape
    n_img = S1 if double else S1 - T

    def proj(name, t):
        return F.linear(t, p[name][0], p[name][1]).view(t.shape[0], t.shape[1], heads, -1)

    q, k, v = proj("q", x), proj("k", x), proj("v", x)
    if "norm_q" in p:
        q, k = hunyuan_rmsnorm(q, p["norm_q"], eps), hunyuan_rmsnorm(k, p["norm_k"], eps)
    if rotary is not None:
        q = torch.cat([hunyuan_apply_rotary_emb(q[:, :n_img], *rotary), q[:, n_img:]], dim=1)
        k = torch.cat([hunyuan_apply_rotary_emb(k[:, :n_img], *rotary), k[:, n_img:]], dim=1)
    if double:
        eq, ek, ev = proj("add_q", encoder_hidden), proj("add_k", encoder_hidden), proj("add_v", encoder_hidden)
        if "norm_added_q" in p:
            eq, ek = hunyuan_rmsnorm(eq, p["norm_added_q"], eps), hunyuan_rmsnorm(ek, p["norm_added_k"], eps)
        q, k, v = torch.cat([q, eq], 1), torch.cat([k, ek], 1), torch.cat([v, ev], 1)
    S = q.shape[1]
    mask = None
    if valid_len is not None:
        idx = torch.arange(S)
        mask = (idx[None, None, :] < valid_len[:, None, None]).expand(B, S, S)[:, None]  # keys; padded rows are discarded
    o = sdpa_blhd(q, k, v, attn_mask=mask).reshape(B, S, -1)
    hs, ehs = o[:, :S - T], o[:, S - T:]
    if "out" in p:
        hs = F.linear(hs, p["out"][0], p["out"][1])
    if "add_out" in p:
        ehs = F.linear(ehs, p["add_out"][0], p["add_out"][1])
    return hs, ehs


def max_rel_err(y: Tensor, ref: Tensor) -> float:
    y, ref = y.detach().double().cpu(), ref.detach().double().cpu()
    return float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def cosine(a: Tensor, b: Tensor) -> float:
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------------------------------
# Approximations the CUDA kernels use inside otherwise exact formulas, restated in float32 numpy so that their error
# bounds are pinned on the CPU (tests/test_oracle_golden.py) — test infrastructure like the rest of this file.
# ---------------------------------------------------------------------------------------------------------------------
def erf_abramowitz_stegun_f32(x):
    """erf(x) by Abramowitz & Stegun 7.1.26 in float32, as csrc/geglu.cu evaluates it (the exact-GELU of lvdm's GEGLU,
    lvdm/modules/attention.py:522-529 -> F.gelu). Published bound: |error| <= 1.5e-7 in exact arithmetic; 5.3e-7 as evaluated here in float32."""
    import numpy as np
    x = np.asarray(x, dtype=np.float32)
    ax = np.abs(x)
    t = (np.float32(1.0) / (np.float32(0.3275911) * ax + np.float32(1.0))).astype(np.float32)
    p = np.float32(1.061405429) * t + np.float32(-1.453152027)
    p = p * t + np.float32(1.421413741)
    p = p * t + np.float32(-0.284496736)
    p = p * t + np.float32(0.254829592)
    e = np.float32(1.0) - p * t * np.exp(-(ax * ax), dtype=np.float32)
    return np.copysign(e, x).astype(np.float32)


def gelu_and_grad_f32(g):
    """(gelu(g), d gelu / dg) from the approximation above, as geglu.cu's gelu_f / dgelu_f."""
    import numpy as np
    g = np.asarray(g, dtype=np.float32)
    cdf = np.float32(0.5) * (np.float32(1.0) + erf_abramowitz_stegun_f32(g * np.float32(0.70710678118654752)))
    pdf = np.float32(0.3989422804014327) * np.exp(np.float32(-0.5) * g * g, dtype=np.float32)
    return (g * cdf).astype(np.float32), (cdf + g * pdf).astype(np.float32)


def silu_and_grad_by_tanh_f32(z):
    """silu(z) = h + h tanh(h), h = z / 2;  silu'(z) = s (1 + z (1 - s)), s = (1 + tanh(h)) / 2 — the one-transcendental forms
    of csrc/groupnorm*.cu (exact identities; the kernels evaluate tanh with `tanh.approx.f32`)."""
    import numpy as np
    z = np.asarray(z, dtype=np.float32)
    h = np.float32(0.5) * z
    th = np.tanh(h, dtype=np.float32)
    s = np.float32(0.5) * th + np.float32(0.5)
    return (h * th + h).astype(np.float32), (s * (z * (np.float32(1.0) - s) + np.float32(1.0))).astype(np.float32)


def ex2_poly_f32(x):
    """2^x as csrc/sm100_ptx.cuh `ex2_poly2` evaluates it on the FMA pipe (a share of the softmax exponentials of the
    attention kernels): clamp at -125, Cody-Waite split through the 1.5 * 2^23 magic constant, degree-3 polynomial for the
    fractional part, integer part added into the exponent field. Stated max relative error: 7.5e-5."""
    import numpy as np
    x = np.maximum(np.asarray(x, dtype=np.float32), np.float32(-125.0))
    magic = np.float32(12582912.0)
    r = (x + magic).astype(np.float32)
    n = (r - magic).astype(np.float32)
    f = (x - n).astype(np.float32)
    p = f * np.float32(0.0551716648) + np.float32(0.2426111251)
    p = (p * f + np.float32(0.6932609677)).astype(np.float32)
    p = (p * f + np.float32(0.9999280572)).astype(np.float32)
    bits = p.view(np.int32) + (r.view(np.int32) << 23)
    return bits.astype(np.int32).view(np.float32)
