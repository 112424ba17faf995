"""Block-level drop-in forwards: the reference's transformer blocks re-expressed over the b200vt kernels.

`functional.py` replaces the attention *functions*; this module replaces the `forward` of the blocks that call them, so
that the memory-bound chain around attention (LayerNorm + adaLN modulate, QK-RMSNorm + RoPE, gated residual,
GroupNorm + SiLU) runs in the fused row kernels instead of the reference's elementwise passes. Every function takes
the reference module as `self` and uses only its own parameters and sub-modules (state-dict keys, LoRA targets and
checkpoints are untouched); `patch.patch_blocks()` installs them on the reference classes. Anything outside the CUDA
path raises `functional.Unsupported`, which the patch layer turns into a call of the original forward.

reference forward                                                          replaced by
  hunyuan MMDoubleStreamBlock.forward   hyvideo_t2v/modules/models.py:132-252   hunyuan_double_block_forward
  hunyuan MMSingleStreamBlock.forward   hyvideo_t2v/modules/models.py:326-393   hunyuan_single_block_forward
  wan WanSelfAttention.forward          wan/wan/modules/model.py:127-156        wan_self_attention_forward
  wan WanT2VCrossAttention.forward      wan/wan/modules/model.py:161-181        wan_t2v_cross_attention_forward
  wan WanI2VCrossAttention.forward      wan/wan/modules/model.py:199-225        wan_i2v_cross_attention_forward
  wan WanAttentionBlock.forward         wan/wan/modules/model.py:274-313        wan_attention_block_forward
  lvdm BasicTransformerBlock._forward   lvdm/modules/attention.py:299-310       lvdm_basic_block_forward
  lvdm SpatialTransformer.forward       lvdm/modules/attention.py:376-392       lvdm_spatial_transformer_forward
  lvdm TemporalTransformer.forward      lvdm/modules/attention.py:475-519       lvdm_temporal_transformer_forward
  lvdm ResBlock._forward                lvdm/modules/networks/openaimodel3d.py:229-255   lvdm_resblock_forward
  lvdm TemporalConvBlock.forward        lvdm/modules/networks/openaimodel3d.py:303-310   lvdm_temporal_conv_block_forward
  diffusers CogVideoXAttnProcessor2_0   [ext, diffusers 0.32.2]                 CogVideoXAttnProcessor
  diffusers CogVideoXBlock.forward      [ext, diffusers 0.32.2]                 cogvideox_block_forward
  diffusers HunyuanVideoAttnProcessor2_0 [ext, diffusers 0.32.2]                HunyuanVideoAttnProcessor
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import functional as Fn
from . import sp as _sp
from .functional import Unsupported, _require

_BF16 = torch.bfloat16


def _is_identity(m) -> bool:
    return isinstance(m, nn.Identity)


def _bf16_consumer(*linears) -> None:
    """The row kernels emit bf16 for the Linear that follows. That is what the reference computes under bf16 autocast or
    with bf16 weights; an fp32 model run WITHOUT autocast (plain fp32 inference) must stay on the reference path — handing
    its fp32 Linear a bf16 tensor raises a dtype error instead of falling back."""
    if torch.is_autocast_enabled():
        return
    for lin in linears:
        w = getattr(lin, "weight", None)
        _require(w is None or w.dtype == _BF16, "fp32 weights without autocast stay on the reference path")


def _rms_weight(norm) -> Tuple[Optional[Tensor], float]:
    """(weight, eps) of a reference RMSNorm (hunyuan norm_layers.py:5-59, wan model.py:70-86); Identity -> (None, 0)."""
    if _is_identity(norm):
        return None, 0.0
    _require(hasattr(norm, "weight") and hasattr(norm, "eps") and not isinstance(norm, nn.LayerNorm),
             "only RMSNorm q/k normalisation is on the CUDA path")
    return norm.weight, float(norm.eps)


# =====================================================================================================================
# HunyuanVideo
# =====================================================================================================================
def _hy_qk(x: Tensor, norm, cos: Optional[Tensor], sin: Optional[Tensor]) -> Tensor:
    """RMSNorm over each head + RoPE on the first cos.shape[0] tokens of a (B, L, H, D) view of the fused QKV output."""
    w, eps = _rms_weight(norm)
    if w is None and cos is None:
        return x
    return Fn.qk_rmsnorm_rope(x, w, cos, sin, per_head=True, eps=eps if w is not None else 1e-6)


def hunyuan_parallel_attention(hybrid_seq_parallel_attn, q, k, v, img_q_len, img_kv_len, cu_seqlens_q, cu_seqlens_kv):
    """Same signature and result as the reference `parallel_attention` (attenion.py:159-212): sequence-parallel attention
    of the image shard + replicated valid text ("rear" joint strategy) through `hybrid_seq_parallel_attn`, and a local
    attention of the padding tail among itself, concatenated to (B, S, H*D). Like the reference it reads
    cu_seqlens[1] on the host (one sync per call) and assumes batch 1."""
    n_q, n_k = int(cu_seqlens_q[1]), int(cu_seqlens_kv[1])
    attn1 = hybrid_seq_parallel_attn(
        None, q[:, :img_q_len], k[:, :img_kv_len], v[:, :img_kv_len], dropout_p=0.0, causal=False,
        joint_tensor_query=q[:, img_q_len:n_q], joint_tensor_key=k[:, img_kv_len:n_k],
        joint_tensor_value=v[:, img_kv_len:n_k], joint_strategy="rear")
    parts = [attn1]
    if q.shape[1] > n_q:
        # the padding tail attends among itself, locally (the reference calls flash-attn directly, attenion.py:182-197);
        # an injected attention core (CPU tests of the exchange logic) serves the tail as well
        core = getattr(hybrid_seq_parallel_attn, "attn_fn", None)
        if core is None or core is _sp._default_attn:
            parts.append(Fn.attention_blhd(q[:, n_q:], k[:, n_k:], v[:, n_k:]))
        else:
            parts.append(core(q[:, n_q:], k[:, n_k:], v[:, n_k:], None))
    attn = torch.cat(parts, dim=1) if len(parts) > 1 else attn1
    b, s, a, d = attn.shape
    return attn.reshape(b, s, a * d)


def _hy_attention(self, q, k, v, img_len, cu_seqlens_q, cu_seqlens_kv, max_seqlen_q, max_seqlen_kv, batch):
    if not getattr(self, "hybrid_seq_parallel_attn", None):
        return Fn.hunyuan_attention(q, k, v, cu_seqlens_q=cu_seqlens_q, cu_seqlens_kv=cu_seqlens_kv,
                                    max_seqlen_q=max_seqlen_q, max_seqlen_kv=max_seqlen_kv, batch_size=batch)
    return hunyuan_parallel_attention(self.hybrid_seq_parallel_attn, q, k, v, img_q_len=img_len, img_kv_len=img_len,
                                      cu_seqlens_q=cu_seqlens_q, cu_seqlens_kv=cu_seqlens_kv)


def _token_replace(condition_type, token_replace_vec, frist_frame_token_num) -> int:
    """Number of leading image rows modulated with the token_replace vectors (hyvideo_i2v/modules/models.py:150-152:
    only condition_type == "token_replace" switches the i2v twins away from the t2v arithmetic)."""
    if condition_type != "token_replace":
        return 0
    _require(token_replace_vec is not None and frist_frame_token_num is not None and int(frist_frame_token_num) > 0,
             "token_replace needs token_replace_vec and a positive frist_frame_token_num")
    return int(frist_frame_token_num)


def _linear2_split(linear2: nn.Module, attn: Tensor, mlp_act: Tensor) -> Tensor:
    """linear2(cat(attn, mlp_act)) (models.py:390) without the concatenation: the weight is applied as two column
    blocks. Anything but a plain nn.Linear (e.g. a LoRA wrapper around linear2) keeps the module call."""
    if type(linear2) is not nn.Linear:
        return linear2(torch.cat((attn, mlp_act), 2))
    c = attn.shape[-1]
    w = linear2.weight
    out = torch.nn.functional.linear(attn, w[:, :c], linear2.bias)
    B, S, _ = out.shape
    return torch.addmm(out.view(B * S, -1), mlp_act.reshape(B * S, -1), w[:, c:].t()).view(B, S, -1)


def hunyuan_double_block_forward(self, img: Tensor, txt: Tensor, vec: Tensor, cu_seqlens_q: Optional[Tensor] = None,
                                 cu_seqlens_kv: Optional[Tensor] = None, max_seqlen_q: Optional[int] = None,
                                 max_seqlen_kv: Optional[int] = None, freqs_cis: tuple = None,
                                 condition_type: Optional[str] = None, token_replace_vec: Optional[Tensor] = None,
                                 frist_frame_token_num: Optional[int] = None):
    """Drop-in body of MMDoubleStreamBlock.forward — the t2v signature (hyvideo_t2v/modules/models.py:132-143) and the
    i2v twin's, which the new-style flow uses for T2V and I2V alike and always calls with 11 positionals
    (hyvideo_i2v/modules/models.py:136-149, 749-761). Per stream: one LayerNorm+modulate pass, the module's own fused QKV
    Linear, per-head RMSNorm (+RoPE) of q and k written directly into the joint [img; txt] tensors (no fp32 temporaries,
    no rotate_half copies, no torch.cat), joint varlen attention, and the two gated residuals as single passes.
    condition_type == "token_replace": the first frist_frame_token_num image rows take the modulation vectors computed
    from token_replace_vec (modulate_layers.py:37-96 of the i2v package)."""
    _require(img.is_cuda and img.dtype == _BF16 and txt.dtype == _BF16, "bf16 CUDA activations only")
    _require(cu_seqlens_q is not None and cu_seqlens_kv is not None, "the block needs cu_seqlens (as the reference asserts)")
    B, L, C = img.shape
    T = txt.shape[1]
    H = self.heads_num
    D = C // H
    _require(D in (64, 128), f"head dim {D} stays on the reference path")
    assert cu_seqlens_q.shape[0] == 2 * B + 1, f"cu_seqlens_q.shape:{cu_seqlens_q.shape}, img.shape[0]:{B}"
    cos, sin = freqs_cis if freqs_cis is not None else (None, None)
    ff = _token_replace(condition_type, token_replace_vec, frist_frame_token_num)
    _bf16_consumer(self.img_attn_qkv, self.txt_attn_qkv)

    i_sh1, i_sc1, i_g1, i_sh2, i_sc2, i_g2 = self.img_mod(vec).chunk(6, dim=-1)
    t_sh1, t_sc1, t_g1, t_sh2, t_sc2, t_g2 = self.txt_mod(vec).chunk(6, dim=-1)
    r_sh1 = r_sc1 = r_g1 = r_sh2 = r_sc2 = r_g2 = None
    if ff:
        r_sh1, r_sc1, r_g1, r_sh2, r_sc2, r_g2 = self.img_mod(token_replace_vec).chunk(6, dim=-1)

    img_qkv = self.img_attn_qkv(Fn.ln_modulate(img, i_sh1, i_sc1, eps=self.img_norm1.eps, tr_shift=r_sh1, tr_scale=r_sc1,
                                               first_frame_tokens=ff)).view(B, L, 3, H, D)
    txt_qkv = self.txt_attn_qkv(Fn.ln_modulate(txt, shift=t_sh1, scale=t_sc1, eps=self.txt_norm1.eps)).view(B, T, 3, H, D)
    q, k, v = Fn.hunyuan_joint_qkv(img_qkv, txt_qkv, self.img_attn_q_norm, self.img_attn_k_norm, self.txt_attn_q_norm,
                                   self.txt_attn_k_norm, cos, sin)

    attn = _hy_attention(self, q, k, v, L, cu_seqlens_q, cu_seqlens_kv, max_seqlen_q, max_seqlen_kv, B)
    img_attn, txt_attn = attn[:, :L], attn[:, L:]

    img = Fn.gate_residual(img, self.img_attn_proj(img_attn), i_g1, tr_gate=r_g1, first_frame_tokens=ff)
    img = Fn.gate_residual(img, self.img_mlp(Fn.ln_modulate(img, i_sh2, i_sc2, eps=self.img_norm2.eps, tr_shift=r_sh2,
                                                            tr_scale=r_sc2, first_frame_tokens=ff)),
                           i_g2, tr_gate=r_g2, first_frame_tokens=ff)
    txt = Fn.gate_residual(txt, self.txt_attn_proj(txt_attn), t_g1)
    txt = Fn.gate_residual(txt, self.txt_mlp(Fn.ln_modulate(txt, shift=t_sh2, scale=t_sc2, eps=self.txt_norm2.eps)), t_g2)
    return img, txt


def hunyuan_single_block_forward(self, x: Tensor, vec: Tensor, txt_len: int, cu_seqlens_q: Optional[Tensor] = None,
                                 cu_seqlens_kv: Optional[Tensor] = None, max_seqlen_q: Optional[int] = None,
                                 max_seqlen_kv: Optional[int] = None, freqs_cis: Tuple[Tensor, Tensor] = None,
                                 condition_type: Optional[str] = None, token_replace_vec: Optional[Tensor] = None,
                                 frist_frame_token_num: Optional[int] = None) -> Tensor:
    """Drop-in body of MMSingleStreamBlock.forward (t2v models.py:326-393; i2v twin :371-462 with its three extra
    arguments): RoPE covers the first S - txt_len tokens only, which the fused kernel expresses through the table length,
    so the reference's split / rotate / cat of q and k disappears; linear2 runs on its two operands without the
    concatenation."""
    _require(x.is_cuda and x.dtype == _BF16, "bf16 CUDA activations only")
    _require(cu_seqlens_q is not None and cu_seqlens_kv is not None, "the block needs cu_seqlens (as the reference asserts)")
    B, S, C = x.shape
    H = self.heads_num
    D = C // H
    _require(D in (64, 128), f"head dim {D} stays on the reference path")
    assert cu_seqlens_q.shape[0] == 2 * B + 1, f"cu_seqlens_q.shape:{cu_seqlens_q.shape}, x.shape[0]:{B}"
    cos = sin = None
    if freqs_cis is not None and freqs_cis[0] is not None:
        cos, sin = freqs_cis
        _require(cos.shape[0] == S - txt_len, "RoPE table must cover exactly the image tokens")
    ff = _token_replace(condition_type, token_replace_vec, frist_frame_token_num)
    _bf16_consumer(self.linear1)

    sh, sc, gate = self.modulation(vec).chunk(3, dim=-1)
    r_sh = r_sc = r_gate = None
    if ff:
        r_sh, r_sc, r_gate = self.modulation(token_replace_vec).chunk(3, dim=-1)
    lin = self.linear1(Fn.ln_modulate(x, sh, sc, eps=self.pre_norm.eps, tr_shift=r_sh, tr_scale=r_sc,
                                      first_frame_tokens=ff))
    qkv = lin[..., : 3 * C].unflatten(-1, (3, H, D))
    mlp = lin[..., 3 * C:]
    q = _hy_qk(qkv[:, :, 0], self.q_norm, cos, sin)
    k = _hy_qk(qkv[:, :, 1], self.k_norm, cos, sin)
    v = qkv[:, :, 2]
    attn = _hy_attention(self, q, k, v, S - txt_len, cu_seqlens_q, cu_seqlens_kv, max_seqlen_q, max_seqlen_kv, B)
    out = _linear2_split(self.linear2, attn, self.mlp_act(mlp))
    return Fn.gate_residual(x, out, gate, tr_gate=r_gate, first_frame_tokens=ff)


# =====================================================================================================================
# Wan2.1
# =====================================================================================================================
_WAN_ROPE_CACHE: dict = {}


def wan_rope_tables(grid_size, freqs: Tensor, device) -> Tuple[Tensor, Tensor]:
    """cos/sin tables (f*h*w, D) fp32 in the interleaved (repeat_interleave(2)) form the fused kernel takes, equal to the
    per-token complex factors rope_apply multiplies by (model.py:40-67): the (1024, D/2) complex table is split into
    [D/2 - 2(D/6), D/6, D/6] columns indexed by frame, row and column of the token."""
    f, h, w = (int(v) for v in grid_size)
    key = (f, h, w, freqs.data_ptr(), str(device))
    hit = _WAN_ROPE_CACHE.get(key)
    if hit is not None:
        return hit
    c = freqs.shape[1]
    fs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    fr = torch.cat([fs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), fs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                    fs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, -1)
    out = (fr.real.float().repeat_interleave(2, dim=1).contiguous().to(device),
           fr.imag.float().repeat_interleave(2, dim=1).contiguous().to(device))
    if len(_WAN_ROPE_CACHE) > 16:
        _WAN_ROPE_CACHE.clear()
    _WAN_ROPE_CACHE[key] = out
    return out


def _wan_half(t: Tensor) -> Tensor:
    _require(t.dtype in (_BF16, torch.float32), "Wan projections must produce bf16 (autocast) or fp32 tensors")
    return t if t.dtype == _BF16 else t.to(_BF16)


def _wan_norm_rope(x4: Tensor, norm, grid_sizes, freqs) -> Tensor:
    """WanRMSNorm over the whole token (dim = H*D) + RoPE, one pass. x4 (B, L, H, D) bf16."""
    w, eps = _rms_weight(norm)
    if grid_sizes is None:
        return x4 if w is None else Fn.qk_rmsnorm_rope(x4, w, None, None, per_head=False, eps=eps)
    grids = grid_sizes.tolist() if isinstance(grid_sizes, Tensor) else [list(g) for g in grid_sizes]
    eps = eps if w is not None else 1e-6
    if all(g == grids[0] for g in grids):
        cos, sin = wan_rope_tables(grids[0], freqs, x4.device)
        return Fn.qk_rmsnorm_rope(x4, w, cos, sin, per_head=False, eps=eps)
    outs = []
    for i, g in enumerate(grids):  # per-sample grids (the reference loops over samples too, model.py:47-64)
        cos, sin = wan_rope_tables(g, freqs, x4.device)
        outs.append(Fn.qk_rmsnorm_rope(x4[i:i + 1], w, cos, sin, per_head=False, eps=eps))
    return torch.cat(outs, dim=0)


def wan_self_attention_forward(self, x: Tensor, seq_lens: Tensor, grid_sizes: Tensor, freqs: Tensor) -> Tensor:
    """Drop-in body of WanSelfAttention.forward: q/k/v/o stay the module's Linears; RMSNorm(dim) + RoPE is one bf16 pass
    per tensor instead of an fp32 norm and a float64 complex multiply."""
    _require(x.is_cuda, "CUDA activations only")
    _require(tuple(self.window_size) == (-1, -1), "windowed attention stays on the reference path")
    b, s, n, d = *x.shape[:2], self.num_heads, self.head_dim
    _require(d in (64, 128), f"head dim {d} stays on the reference path")
    q = _wan_norm_rope(_wan_half(self.q(x)).view(b, s, n, d), self.norm_q, grid_sizes, freqs)
    k = _wan_norm_rope(_wan_half(self.k(x)).view(b, s, n, d), self.norm_k, grid_sizes, freqs)
    v = _wan_half(self.v(x)).view(b, s, n, d)
    out = Fn.wan_flash_attention(q, k, v, k_lens=seq_lens, window_size=self.window_size)
    return self.o(out.flatten(2))


def wan_t2v_cross_attention_forward(self, x: Tensor, context: Tensor, context_lens: Optional[Tensor]) -> Tensor:
    """Drop-in body of WanT2VCrossAttention.forward."""
    _require(x.is_cuda, "CUDA activations only")
    b, n, d = x.size(0), self.num_heads, self.head_dim
    _require(d in (64, 128), f"head dim {d} stays on the reference path")
    q = _wan_norm_rope(_wan_half(self.q(x)).view(b, -1, n, d), self.norm_q, None, None)
    k = _wan_norm_rope(_wan_half(self.k(context)).view(b, -1, n, d), self.norm_k, None, None)
    v = _wan_half(self.v(context)).view(b, -1, n, d)
    out = Fn.wan_flash_attention(q, k, v, k_lens=context_lens)
    return self.o(out.flatten(2))


def wan_i2v_cross_attention_forward(self, x: Tensor, context: Tensor, context_lens: Optional[Tensor]) -> Tensor:
    """Drop-in body of WanI2VCrossAttention.forward: the first 257 context tokens are CLIP image tokens with their own
    k/v projections; the two attention results are summed before the output projection."""
    _require(x.is_cuda, "CUDA activations only")
    context_img, context = context[:, :257], context[:, 257:]
    b, n, d = x.size(0), self.num_heads, self.head_dim
    _require(d in (64, 128), f"head dim {d} stays on the reference path")
    q = _wan_norm_rope(_wan_half(self.q(x)).view(b, -1, n, d), self.norm_q, None, None)
    k = _wan_norm_rope(_wan_half(self.k(context)).view(b, -1, n, d), self.norm_k, None, None)
    v = _wan_half(self.v(context)).view(b, -1, n, d)
    k_img = _wan_norm_rope(_wan_half(self.k_img(context_img)).view(b, -1, n, d), self.norm_k_img, None, None)
    v_img = _wan_half(self.v_img(context_img)).view(b, -1, n, d)
    img_x = Fn.wan_flash_attention(q, k_img, v_img, k_lens=None)
    out = Fn.wan_flash_attention(q, k, v, k_lens=context_lens)
    return self.o(out.flatten(2) + img_x.flatten(2))


def _wan_ln(norm, x: Tensor) -> Tensor:
    """WanLayerNorm (fp32 statistics, optional affine, model.py:89-99) or Identity -> bf16 for the following Linear."""
    if _is_identity(norm):
        return x
    return Fn.ln_modulate(x, None, None, getattr(norm, "weight", None), getattr(norm, "bias", None), norm.eps)


def wan_attention_block_forward(self, x: Tensor, e: Tensor, seq_lens, grid_sizes, freqs, context, context_lens) -> Tensor:
    """Drop-in body of WanAttentionBlock.forward. The residual stream keeps the dtype it arrives in (fp32 in the
    reference pipeline: `x + y * e[2]` promotes it); every LayerNorm+modulate reads it once and emits the bf16 tensor
    the next Linear consumes, every gated residual is one read-modify-write pass."""
    _require(x.is_cuda and x.dtype in (_BF16, torch.float32) and x.dim() == 3, "CUDA bf16/fp32 (B,L,C) activations only")
    assert e.dtype == torch.float32
    _bf16_consumer(self.self_attn.q, self.ffn[0])
    e6 = (self.modulation.float() + e).chunk(6, dim=1)  # six (B, 1, C) fp32 tensors

    y = self.self_attn(Fn.ln_modulate(x, shift=e6[0], scale=e6[1], eps=self.norm1.eps), seq_lens, grid_sizes, freqs)
    x = Fn.gate_residual(x, _wan_half(y), e6[2])
    x = Fn.gate_residual(x, _wan_half(self.cross_attn(_wan_ln(self.norm3, x), context, context_lens)), None)
    y = self.ffn(Fn.ln_modulate(x, shift=e6[3], scale=e6[4], eps=self.norm2.eps))
    return Fn.gate_residual(x, _wan_half(y), e6[5])


# =====================================================================================================================
# lvdm (VideoCrafter / DynamiCrafter 3D-UNet)
# =====================================================================================================================
def _lvdm_ln(norm: nn.LayerNorm, x: Tensor) -> Tensor:
    _require(x.is_cuda and x.dtype in (_BF16, torch.float32) and x.shape[-1] % 8 == 0, "CUDA bf16/fp32 activations only")
    _require(x.dtype == _BF16 or torch.is_autocast_enabled() or norm.weight.dtype == _BF16,
             "fp32 activations with fp32 weights and no autocast stay on the reference path")
    return Fn.layer_norm(x, norm.weight, norm.bias, norm.eps)


def lvdm_basic_block_forward(self, x: Tensor, context: Optional[Tensor] = None, mask: Optional[Tensor] = None) -> Tensor:
    """Drop-in body of BasicTransformerBlock._forward (the checkpoint wrapper in .forward stays the reference's):
    the three affine LayerNorms run in the row kernel; attn1/attn2 are the module's CrossAttention (whose forward
    patch_lvdm replaces); the GEGLU feed-forward stays cuBLAS."""
    x = self.attn1(_lvdm_ln(self.norm1, x), context=context if self.disable_self_attn else None, mask=mask) + x
    x = self.attn2(_lvdm_ln(self.norm2, x), context=context, mask=mask) + x
    x = self.ff(_lvdm_ln(self.norm3, x)) + x
    return x


def _lvdm_gn(norm: nn.GroupNorm, x: Tensor, silu: bool = False, addend: Optional[Tensor] = None) -> Tensor:
    _require(x.is_cuda and x.dtype in (_BF16, torch.float32), "CUDA bf16/fp32 activations only")
    _require(norm.affine, "GroupNorm without affine parameters stays on the reference path")
    return Fn.groupnorm_silu(x, norm.weight, norm.bias, norm.num_groups, norm.eps, silu=silu, addend=addend)


_FOLD_BIAS = os.environ.get("B200VT_GN_FOLD_BIAS", "1") != "0"


def _conv_bias_deferred(conv: nn.Module, x: Tensor):
    """conv(x) WITHOUT its bias and the bias to fold into the next GroupNorm (fp32 (C,)), or (conv(x), None) when the bias
    cannot be deferred: not a plain Conv2d / Conv3d, no bias, a trainable bias (its gradient would be lost), an activation
    that is not channels-last (the NCHW GroupNorm kernels take no addend), or B200VT_GN_FOLD_BIAS=0. PyTorch adds a cuDNN
    convolution's bias in a separate broadcast kernel over the whole activation; deferring it removes that pass."""
    if (_FOLD_BIAS and type(conv) in (nn.Conv2d, nn.Conv3d) and conv.bias is not None and not conv.bias.requires_grad
            and _is_cl(x) and x.dtype == _BF16 and conv.out_channels % 8 == 0 and conv.out_channels <= 4096):
        y = conv._conv_forward(x, conv.weight, None)
        if _is_cl(y):
            return y, conv.bias
        return y + conv.bias.view(1, -1, *([1] * (y.dim() - 2))).to(y.dtype), None
    return conv(x), None


def _is_cl(x: Tensor) -> bool:
    """Channels-last activation: (N, C, *spatial) whose memory is (N, *spatial, C) (torch.channels_last / channels_last_3d)."""
    from .ops import _channels_last
    return x.dim() >= 4 and _channels_last(x)


def _cl_weights(conv) -> bool:
    """The module opted into the channels-last flow: its convolution weight was converted once by
    patch.lvdm_channels_last(model) (otherwise cuDNN would re-lay the filter on every call)."""
    w = getattr(conv, "weight", None)
    return w is not None and w.dim() >= 4 and not w.is_contiguous() and w.stride(1) == 1


def lvdm_spatial_transformer_forward(self, x: Tensor, context: Optional[Tensor] = None) -> Tensor:
    """Drop-in body of SpatialTransformer.forward: GroupNorm(32) with fp32 statistics in one kernel, then the
    reference's own projections and blocks. With a channels-last activation (the layout the UNet's convolutions run in
    after patch.lvdm_channels_last) both layout changes of the reference — `b c h w -> b (h w) c` before proj_in and its
    inverse after proj_out (attention.py:381, 389) — are views: GroupNorm reads and writes (b, h*w, c) memory directly."""
    b, c, h, w = x.shape
    x_in = x
    x = _lvdm_gn(self.norm, x)
    if self.use_linear and _is_cl(x):
        x = self.proj_in(x.permute(0, 2, 3, 1).reshape(b, h * w, c))  # free view of the channels-last tensor
        for block in self.transformer_blocks:
            x = block(x, context=context)
        x = self.proj_out(x)
        return x.view(b, h, w, -1).permute(0, 3, 1, 2) + x_in       # a channels-last (b, c, h, w) view: no copy
    if not self.use_linear:
        x = self.proj_in(x)
    x = x.flatten(2).transpose(1, 2).contiguous()  # b c h w -> b (h w) c
    if self.use_linear:
        x = self.proj_in(x)
    for block in self.transformer_blocks:
        x = block(x, context=context)
    if self.use_linear:
        x = self.proj_out(x)
    x = x.transpose(1, 2).reshape(b, -1, h, w).contiguous()  # b (h w) c -> b c h w
    if not self.use_linear:
        x = self.proj_out(x)
    return x + x_in


def lvdm_temporal_transformer_forward(self, x: Tensor, context: Optional[Tensor] = None) -> Tensor:
    """Drop-in body of TemporalTransformer.forward for the configurations the reference trains (only_self_att, with or
    without the causal mask); the per-sample cross-attention loop (:499-509) stays on the reference path. A channels-last
    (channels_last_3d) input keeps its layout: GroupNorm runs on (b, t*h*w, c) memory and the result of the block is
    written back in that layout, so the surrounding `(b f) c h w <-> b c f h w` rearranges stay views."""
    _require(self.only_self_att, "temporal cross-attention stays on the reference path")
    b, c, t, h, w = x.shape
    x_in = x
    cl = _is_cl(x)
    x = _lvdm_gn(self.norm, x)
    if self.use_linear:
        x = x.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c)  # b c t h w -> (b h w) t c   (one copy)
        x = self.proj_in(x)
    else:
        x = x.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, t)  # b c t h w -> (b h w) c t
        x = self.proj_in(x).transpose(1, 2).contiguous()
    mask = None
    if self.causal_attention:
        mask = self.mask.to(x.device).expand(b * h * w, -1, -1)
    for block in self.transformer_blocks:
        x = block(x, mask=mask)
    fmt = torch.channels_last_3d if cl else torch.contiguous_format
    if self.use_linear:
        x = self.proj_out(x)
        x = x.view(b, h, w, t, -1).permute(0, 4, 3, 1, 2).contiguous(memory_format=fmt)  # (b h w) t c -> b c t h w
    else:
        x = self.proj_out(x.transpose(1, 2).contiguous())
        x = x.view(b, h, w, -1, t).permute(0, 3, 4, 1, 2).contiguous(memory_format=fmt)
    return x + x_in


def lvdm_resblock_forward(self, x: Tensor, emb: Tensor, batch_size: Optional[int] = None) -> Tensor:
    """Drop-in body of ResBlock._forward: both GroupNormSpecific + SiLU pairs (fp32 statistics, utils.py:192-203) are one
    kernel each; convolutions, up/down-sampling and the temporal conv block stay the module's own layers."""
    _require(x.is_cuda and x.dtype in (_BF16, torch.float32), "CUDA bf16/fp32 activations only")
    in_norm, in_conv = self.in_layers[0], self.in_layers[-1]
    if x.dim() == 4 and _cl_weights(in_conv) and not _is_cl(x):
        x = x.contiguous(memory_format=torch.channels_last)  # enter the channels-last flow (patch.lvdm_channels_last)
    h = _lvdm_gn(in_norm, x, silu=True)
    if self.updown:
        h = self.h_upd(h)
        x = self.x_upd(x)
    emb_out = self.emb_layers(emb)
    out_norm, out_rest = self.out_layers[0], self.out_layers[2:]  # [norm, SiLU, Dropout, conv]
    fold = not self.use_scale_shift_norm and emb_out.dim() == 2 and not (emb_out.requires_grad and torch.is_grad_enabled())
    h, bias = _conv_bias_deferred(in_conv, h) if fold else (in_conv(h), None)
    if bias is not None:
        # GroupNorm(conv(h) + bias + emb_out[..., None, None]) with both broadcast terms folded into the kernel's constants
        h = out_rest(_lvdm_gn(out_norm, h, silu=True, addend=emb_out.float() + bias.float()))
    else:
        emb_out = emb_out.type(h.dtype)
        while len(emb_out.shape) < len(h.shape):
            emb_out = emb_out[..., None]
        if self.use_scale_shift_norm:
            scale, shift = torch.chunk(emb_out, 2, dim=1)
            h = _lvdm_gn(out_norm, h) * (1 + scale) + shift
            h = self.out_layers[1:](h)
        else:
            h = out_rest(_lvdm_gn(out_norm, h + emb_out, silu=True))
    h = self.skip_connection(x) + h
    if self.use_temporal_conv and batch_size:
        bt, ch, hh, ww = h.shape
        h = h.view(batch_size, bt // batch_size, ch, hh, ww).transpose(1, 2)  # (b t) c h w -> b c t h w
        h = self.temopral_conv(h)
        h = h.transpose(1, 2).reshape(bt, ch, hh, ww)
    return h


def lvdm_temporal_conv_block_forward(self, x: Tensor) -> Tensor:
    """Drop-in body of TemporalConvBlock.forward (openaimodel3d.py:258-310, one per ResBlock with use_temporal_conv: 22 in
    the VideoCrafter2 UNet): four stages of GroupNorm(32) -> SiLU -> [Dropout] -> Conv3d on (b, c, t, h, w) plus the
    identity. Each GroupNorm + SiLU pair is one pass of the 5-D cluster kernel; the (3,1,1) convolutions stay cuDNN."""
    _require(x.is_cuda and x.dtype in (_BF16, torch.float32) and x.dim() == 5, "CUDA bf16/fp32 (b, c, t, h, w) activations only")
    h, bias = x, None
    stages = (self.conv1, self.conv2, self.conv3, self.conv4)
    for i, stage in enumerate(stages):
        norm, act = stage[0], stage[1]
        _require(isinstance(norm, nn.GroupNorm) and isinstance(act, nn.SiLU), "unexpected TemporalConvBlock layout")
        h = _lvdm_gn(norm, h, silu=True, addend=bias)  # the previous convolution's bias rides in this GroupNorm
        bias = None
        if i + 1 < len(stages):
            h = stage[2:-1](h)  # [Dropout]
            h, bias = _conv_bias_deferred(stage[-1], h)
        else:
            h = stage[2:](h)
    return h + x


# =====================================================================================================================
# diffusers attention processors (CogVideoX, diffusers-HunyuanVideo) — duck-typed, no diffusers import
# =====================================================================================================================
def _proc_qk_norm(norm, x4: Tensor) -> Tensor:
    """diffusers `Attention.norm_q / norm_k` applied per head on (B, L, H, D): LayerNorm(D) for CogVideoX (qk_norm=
    "layer_norm"), RMSNorm(D) for HunyuanVideo (qk_norm="rms_norm"); None -> unchanged."""
    if norm is None:
        return x4
    if isinstance(norm, nn.LayerNorm):
        B, L, H, D = x4.shape
        return Fn.layer_norm(x4.reshape(B, L * H, D), norm.weight, norm.bias, norm.eps).view(B, L, H, D)
    w, eps = _rms_weight(norm)
    return Fn.qk_rmsnorm_rope(x4, w, None, None, per_head=True, eps=eps)


def _proc_rope(x4: Tensor, image_rotary_emb, start: int, length: int) -> Tensor:
    """diffusers apply_rotary_emb(use_real=True, unbind_dim=-1) — the same interleaved-pair rotation as hunyuan's —
    on tokens [start, start + length) of (B, L, H, D)."""
    if image_rotary_emb is None:
        return x4
    cos, sin = image_rotary_emb
    if start == 0:
        _require(cos.shape[0] == length, "rotary table must cover exactly the rotated tokens")
        return Fn.qk_rmsnorm_rope(x4, None, cos, sin)
    head, tail = x4[:, :start], x4[:, start:]
    return torch.cat([head, Fn.qk_rmsnorm_rope(tail, None, cos, sin)], dim=1)


class CogVideoXAttnProcessor:
    """`attn.set_processor(CogVideoXAttnProcessor())` — protocol of diffusers 0.32.2 `CogVideoXAttnProcessor2_0`
    (reference call sites: cogvideo_hf/cogvideo_pl.py:123, 862-868; semantics cross-checked against the in-tree SAT
    description, cogvideo_sat/dit_video_concat.py:263-427): text tokens first, q/k/v Linears on the concatenation,
    per-head LayerNorm on q and k, RoPE on the video tokens only, dense joint attention, output projection, split."""

    def __call__(self, attn, hidden_states: Tensor, encoder_hidden_states: Tensor, attention_mask: Optional[Tensor] = None,
                 image_rotary_emb=None) -> Tuple[Tensor, Tensor]:
        _require(attention_mask is None, "attention masks stay on the stock processor")
        text_len = encoder_hidden_states.size(1)
        x = torch.cat([encoder_hidden_states, hidden_states], dim=1)
        _require(x.is_cuda and x.dtype == _BF16, "bf16 CUDA activations only")
        B, S, _ = x.shape
        H = attn.heads
        q, k, v = attn.to_q(x), attn.to_k(x), attn.to_v(x)
        D = q.shape[-1] // H
        _require(D in (64, 128), f"head dim {D} stays on the stock processor")
        q = _proc_qk_norm(getattr(attn, "norm_q", None), q.view(B, S, H, D))
        k = _proc_qk_norm(getattr(attn, "norm_k", None), k.view(B, S, H, D))
        if image_rotary_emb is not None:
            q = _proc_rope(q, image_rotary_emb, text_len, S - text_len)
            if not getattr(attn, "is_cross_attention", False):
                k = _proc_rope(k, image_rotary_emb, text_len, S - text_len)
        out = Fn.attention_blhd(q, k, v.view(B, S, H, D)).reshape(B, S, H * D)
        out = attn.to_out[1](attn.to_out[0](out))
        return out[:, text_len:], out[:, :text_len]


def _cog_norm_zero(norm, hidden: Tensor, encoder: Tensor, temb: Tensor):
    """diffusers `CogVideoXLayerNormZero.forward`: six modulation vectors from Linear(SiLU(temb)); both streams go through
    the same affine LayerNorm and their own (1 + scale) / shift; returns the two gates as (B, C)."""
    shift, scale, gate, e_shift, e_scale, e_gate = norm.linear(norm.silu(temb)).chunk(6, dim=1)
    ln = norm.norm
    w, b = getattr(ln, "weight", None), getattr(ln, "bias", None)
    return (Fn.ln_modulate(hidden, shift, scale, w, b, ln.eps), Fn.ln_modulate(encoder, e_shift, e_scale, w, b, ln.eps),
            gate, e_gate)


def cogvideox_block_forward(self, hidden_states: Tensor, encoder_hidden_states: Tensor, temb: Tensor,
                            image_rotary_emb=None, attention_kwargs=None) -> Tuple[Tensor, Tensor]:
    """Drop-in body of diffusers 0.32.2 `CogVideoXBlock.forward` (the denoiser block behind the reference's
    cogvideo_hf/cogvideo_pl.py:862-868; parity unpinned): norm1 (LayerNormZero) -> attn1 (the module's own
    `Attention`, i.e. whatever processor is installed — `CogVideoXAttnProcessor` after set_diffusers_processors) -> gated
    residuals -> norm2 -> feed-forward over [text; video] -> gated residuals. The four LayerNorm + modulate passes and the
    four gate-multiply-add passes run in the fused row kernels."""
    _require(hidden_states.is_cuda and hidden_states.dtype == _BF16 and encoder_hidden_states.dtype == _BF16,
             "bf16 CUDA activations only")
    _require(not attention_kwargs, "attention_kwargs stay on the stock block")
    T = encoder_hidden_states.size(1)
    hn, en, gate, e_gate = _cog_norm_zero(self.norm1, hidden_states, encoder_hidden_states, temb)
    ah, ae = self.attn1(hidden_states=hn, encoder_hidden_states=en, image_rotary_emb=image_rotary_emb)
    hidden_states = Fn.gate_residual(hidden_states, ah, gate)
    encoder_hidden_states = Fn.gate_residual(encoder_hidden_states, ae, e_gate)
    hn, en, gate, e_gate = _cog_norm_zero(self.norm2, hidden_states, encoder_hidden_states, temb)
    ff = self.ff(torch.cat([en, hn], dim=1))
    hidden_states = Fn.gate_residual(hidden_states, ff[:, T:], gate)
    encoder_hidden_states = Fn.gate_residual(encoder_hidden_states, ff[:, :T], e_gate)
    return hidden_states, encoder_hidden_states


_KEY_LENS_CACHE: dict = {}


class HunyuanVideoAttnProcessor:
    """Protocol of diffusers 0.32.2 `HunyuanVideoAttnProcessor2_0` (reference call site: hyvideo_t2v/hunyuanvideo.py:209,
    946-955): video tokens first; double-stream blocks (attn.add_q_proj present) project the text stream separately,
    single-stream blocks receive [video; text] already concatenated (encoder_hidden_states is only a length marker);
    per-head RMSNorm on q/k, RoPE on the video tokens, joint attention. The only mask shape on the CUDA path is the
    reference's key-padding mask (B, 1, 1|S, S) or (B, S), turned into per-sample key lengths."""

    @staticmethod
    def _key_lens(attention_mask: Optional[Tensor], B: int, S: int) -> Optional[Tensor]:
        if attention_mask is None:
            return None
        m = attention_mask
        _require(m.dtype == torch.bool, "only boolean key-padding masks are on the CUDA path")
        if m.dim() == 4:
            m = m[:, 0, 0]
        _require(m.shape == (B, S), "unsupported mask shape")
        # The denoiser passes the SAME mask tensor to every block of every step: validate it (one host sync) and turn it
        # into key lengths once per tensor, not once per attention call.
        key = (m.data_ptr(), tuple(m.shape), m._version, str(m.device))
        hit = _KEY_LENS_CACHE.get(key)
        if hit is None:
            lens = m.sum(dim=1).to(torch.int32)
            # a key-padding mask must be a prefix of ones for key lengths to express it
            ok = bool((m == (torch.arange(S, device=m.device)[None] < lens[:, None])).all())
            if len(_KEY_LENS_CACHE) > 64:
                _KEY_LENS_CACHE.clear()
            hit = _KEY_LENS_CACHE[key] = (lens if ok else None, m)  # keep m alive: the key holds its address
        _require(hit[0] is not None, "non-prefix mask")
        return hit[0]

    def __call__(self, attn, hidden_states: Tensor, encoder_hidden_states: Optional[Tensor] = None,
                 attention_mask: Optional[Tensor] = None, image_rotary_emb=None) -> Tuple[Tensor, Optional[Tensor]]:
        double = getattr(attn, "add_q_proj", None) is not None
        if not double and encoder_hidden_states is not None:
            hidden_states = torch.cat([hidden_states, encoder_hidden_states], dim=1)
        _require(hidden_states.is_cuda and hidden_states.dtype == _BF16, "bf16 CUDA activations only")
        B, S1, _ = hidden_states.shape
        H = attn.heads
        T = encoder_hidden_states.shape[1] if encoder_hidden_states is not None else 0
        n_img = S1 if double else S1 - T
        q, k, v = attn.to_q(hidden_states), attn.to_k(hidden_states), attn.to_v(hidden_states)
        D = q.shape[-1] // H
        _require(D in (64, 128), f"head dim {D} stays on the stock processor")
        q = _proc_qk_norm(getattr(attn, "norm_q", None), q.view(B, S1, H, D))
        k = _proc_qk_norm(getattr(attn, "norm_k", None), k.view(B, S1, H, D))
        v = v.view(B, S1, H, D)
        if image_rotary_emb is not None:
            cos, sin = image_rotary_emb
            _require(cos.shape[0] == n_img, "rotary table must cover exactly the video tokens")
            q = Fn.qk_rmsnorm_rope(q, None, cos, sin)
            k = Fn.qk_rmsnorm_rope(k, None, cos, sin)
        if double and encoder_hidden_states is not None:
            eq = attn.add_q_proj(encoder_hidden_states).view(B, T, H, D)
            ek = attn.add_k_proj(encoder_hidden_states).view(B, T, H, D)
            ev = attn.add_v_proj(encoder_hidden_states).view(B, T, H, D)
            eq = _proc_qk_norm(getattr(attn, "norm_added_q", None), eq)
            ek = _proc_qk_norm(getattr(attn, "norm_added_k", None), ek)
            q, k, v = torch.cat([q, eq], 1), torch.cat([k, ek], 1), torch.cat([v, ev], 1)
        S = q.shape[1]
        out = Fn.attention_blhd(q, k, v, k_lens=self._key_lens(attention_mask, B, S)).reshape(B, S, H * D)
        if encoder_hidden_states is None:
            return out, None
        hs, ehs = out[:, :S - T], out[:, S - T:]
        if getattr(attn, "to_out", None) is not None:
            hs = attn.to_out[1](attn.to_out[0](hs))
        if getattr(attn, "to_add_out", None) is not None:
            ehs = attn.to_add_out(ehs)
        return hs, ehs
