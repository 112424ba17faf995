"""Reference-signature functions over the b200vt ops (host-side mirror of the reference's Python plugin points).

Each public function keeps the name, argument meaning and return layout of the reference callable it replaces, so the
parity tests read like calls into VideoTuna. Anything the CUDA path does not implement raises `Unsupported`; the
patch layer (patch.py) catches that and routes the call to the untouched reference function — never to a CPU kernel of
our own.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops


class Unsupported(NotImplementedError):
    """The requested variant is outside the CUDA path; callers fall back to the reference implementation."""


_HALF = (torch.bfloat16,)
_LORA_MERGE = os.environ.get("B200VT_LORA_MERGE", "1") != "0"


def _require(cond: bool, why: str):
    if not cond:
        raise Unsupported(why)


def supported_qkv(q: Tensor, k: Tensor, v: Tensor) -> bool:
    return (q.is_cuda and q.dtype in _HALF and k.dtype == q.dtype and v.dtype == q.dtype
            and q.shape[-1] in (64, 128) and k.shape[-1] == q.shape[-1] and v.shape[-1] == q.shape[-1])


# ---------------------------------------------------------------------------------------------------------------------
# generic (B, L, H, D) attention
# ---------------------------------------------------------------------------------------------------------------------
def attention_blhd(q: Tensor, k: Tensor, v: Tensor, softmax_scale: Optional[float] = None,
                   k_lens: Optional[Tensor] = None, cu_seqlens_q: Optional[Tensor] = None,
                   cu_seqlens_k: Optional[Tensor] = None, max_seqlen_q: Optional[int] = None,
                   max_seqlen_k: Optional[int] = None, return_lse: bool = False):
    """softmax(q k^T * scale) v for q (B,Lq,H,D), k/v (B,Lk,H,D) bf16 CUDA tensors -> (B,Lq,H,D).

    k_lens (B,) int32: keys >= k_lens[b] are masked (fixed mode). cu_seqlens_*: packed varlen mode, B must be 1.
    """
    _require(supported_qkv(q, k, v), f"attention needs CUDA bf16 with head dim 64/128, got {q.dtype} {tuple(q.shape)}")
    _require(q.shape[2] == k.shape[2] == v.shape[2], "grouped-query attention is not on this path")
    scale = 1.0 / math.sqrt(q.shape[-1]) if softmax_scale is None else float(softmax_scale)
    if cu_seqlens_q is not None:
        cu_seqlens_q = cu_seqlens_q.to(device=q.device, dtype=torch.int32)
        cu_seqlens_k = cu_seqlens_q if cu_seqlens_k is None else cu_seqlens_k.to(device=q.device, dtype=torch.int32)
        mq = int(max_seqlen_q) if max_seqlen_q is not None else q.shape[1]
        mk = int(max_seqlen_k) if max_seqlen_k is not None else k.shape[1]
        o, lse = ops.attn_fwd(q, k, v, cu_seqlens_q, cu_seqlens_k, None, mq, mk, scale)
    else:
        if k_lens is not None:
            k_lens = k_lens.to(device=q.device, dtype=torch.int32)
        o, lse = ops.attn_fwd(q, k, v, None, None, k_lens, q.shape[1], k.shape[1], scale)
    return (o, lse) if return_lse else o


# ---------------------------------------------------------------------------------------------------------------------
# host-buffer entry point: attention fwd+bwd on pinned host tensors, pipelined over head groups
# ---------------------------------------------------------------------------------------------------------------------
def copy_head_group(dev_t: Tensor, host_t: Tensor, h0: int, to_device: bool, stream) -> None:
    """One strided DMA per sample between heads [h0, h0 + hg) of a pinned host (B, L, H, D) bf16 tensor and a contiguous
    device (B, L, hg, D) tensor, enqueued on `stream` (cudaMemcpy2DAsync: L rows of hg*D elements, host pitch H*D)."""
    from . import _lib
    B, L, hg, D = dev_t.shape
    es = dev_t.element_size()
    for b in range(B):
        hptr = host_t.data_ptr() + (b * host_t.stride(0) + h0 * host_t.stride(2)) * es
        dptr = dev_t.data_ptr() + b * dev_t.stride(0) * es
        hpitch, dpitch, width = host_t.stride(1) * es, hg * D * es, hg * D * es
        if to_device:
            _lib.call("vt_memcpy2d_async", _lib.vp(dptr), dpitch, _lib.vp(hptr), hpitch, width, L, 1, _lib.vp(stream.cuda_stream))
        else:
            _lib.call("vt_memcpy2d_async", _lib.vp(hptr), hpitch, _lib.vp(dptr), dpitch, width, L, 0, _lib.vp(stream.cuda_stream))


class HostAttention:
    """softmax(q k^T * scale) v forward + backward for q, k, v, dO that live in PINNED HOST memory as (B, L, H, D)
    bf16 — the activation-offload case of long-sequence DiT finetuning (one HunyuanVideo layer's q, k, v at 119 056
    tokens is 2.2 GB). Heads are independent, so the work is cut into head groups and pipelined on three streams:
    group g+1 is copied in (strided cudaMemcpy2DAsync straight out of the (B, L, H, D) layout) while group g runs the
    forward and backward kernels and group g-1's o, dq, dk, dv are copied out; PCIe is full duplex, so apart from the
    first copy-in of the first call and the last copy-out of the last the transfers hide behind the kernels (consecutive
    calls pipeline into one another). Input device buffers are allocated once and reused across calls. Same arithmetic as
    attention_blhd + autograd (tests/test_gpu_attention.py)."""

    def __init__(self, B: int, L: int, H: int, D: int, head_groups: int = 4, device=None):
        _require(H % head_groups == 0, f"head_groups={head_groups} must divide H={H}")
        _require(D in (64, 128), "head dim must be 64 or 128")
        self.B, self.L, self.H, self.D, self.G = B, L, H, D, head_groups
        self.hg = H // head_groups
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        shape = (B, L, self.hg, D)

        def buf():
            return torch.empty(shape, dtype=torch.bfloat16, device=self.dev)

        # two sets of inputs (q, k, v, dO): the copy-in of the next group overlaps the kernels of the current one; outputs
        # are fresh allocations kept alive for the copy-out stream (record_stream)
        self.inp = [[buf() for _ in range(4)] for _ in range(2)]
        self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.ev_in = [torch.cuda.Event() for _ in range(2)]      # inputs of slot landed
        self.ev_free = [torch.cuda.Event() for _ in range(2)]    # compute finished reading slot's inputs
        self.ev_done = [torch.cuda.Event() for _ in range(2)]    # outputs of slot computed
        self.n_groups_done = 0   # running counter: slots alternate ACROSS calls (see __call__)

    def synchronize(self) -> None:
        """Wait until every copy-out enqueued so far has landed in the host buffers."""
        self.s_out.synchronize()

    def __call__(self, q: Tensor, k: Tensor, v: Tensor, dout: Tensor, out: Tensor, dq: Tensor, dk: Tensor, dv: Tensor,
                 softmax_scale: Optional[float] = None):
        """All eight tensors: pinned host (B, L, H, D) bf16 with unit D stride and H stride D. Returns when everything
        has been ENQUEUED. Consecutive calls pipeline into one another — the next call's first copy-in runs under this
        call's last kernels, this call's last copy-out under the next call's first — so inputs must stay untouched, and
        outputs are valid, only after synchronize() (or a device synchronize)."""
        for t in (q, k, v, dout, out, dq, dk, dv):
            _require(not t.is_cuda and t.is_pinned() and t.dtype == torch.bfloat16 and tuple(t.shape) ==
                     (self.B, self.L, self.H, self.D) and t.stride(3) == 1 and t.stride(2) == self.D,
                     "HostAttention needs pinned host (B, L, H, D) bf16 tensors with contiguous heads")
        scale = 1.0 / math.sqrt(self.D) if softmax_scale is None else float(softmax_scale)
        cur = torch.cuda.current_stream(self.dev)
        if self.n_groups_done == 0:
            self.s_in.wait_stream(cur)
            self.s_out.wait_stream(cur)
        for g in range(self.G):
            n = self.n_groups_done
            slot, h0 = n & 1, g * self.hg
            with torch.cuda.stream(self.s_in):
                if n >= 2:
                    self.s_in.wait_event(self.ev_free[slot])  # the compute that last used these buffers is done
                for dev_t, host_t in zip(self.inp[slot], (q, k, v, dout)):
                    copy_head_group(dev_t, host_t, h0, True, self.s_in)
                self.ev_in[slot].record(self.s_in)
            cur.wait_event(self.ev_in[slot])
            qd, kd, vd, dod = self.inp[slot]
            o, lse = ops.attn_fwd(qd, kd, vd, None, None, None, self.L, self.L, scale)
            gq, gk, gv = ops.attn_bwd(dod, qd, kd, vd, o, lse, None, None, None, self.L, self.L, scale)
            self.ev_free[slot].record(cur)
            self.ev_done[slot].record(cur)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_done[slot])
                for dev_t, host_t in zip((o, gq, gk, gv), (out, dq, dk, dv)):
                    dev_t.record_stream(self.s_out)  # returned to the allocator only after the copy-out has run
                    copy_head_group(dev_t, host_t, h0, False, self.s_out)
            self.n_groups_done += 1
        return out, dq, dk, dv


# ---------------------------------------------------------------------------------------------------------------------
# HunyuanVideo: attention(q,k,v,mode,...)  (videotuna/models/hunyuan/hyvideo_t2v/modules/attenion.py:60-156)
# ---------------------------------------------------------------------------------------------------------------------
def hunyuan_attention(q, k, v, mode="flash", drop_rate=0, attn_mask=None, causal=False, cu_seqlens_q=None,
                      cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None, batch_size=1):
    """Same signature and return layout ([b, s, a*d]) as the reference `attention`.

    mode="flash": packed two-segment varlen semantics of flash_attn_varlen_func via cu_seqlens (attenion.py:107-120).
    mode="torch": dense attention; attn_mask / causal / dropout are not on the CUDA path (raise Unsupported).
    mode="vanilla": never on the CUDA path (it applies dropout with train=True unconditionally, attenion.py:148).
    """
    _require(mode in ("flash", "torch"), f"mode {mode!r} stays on the reference path")
    _require(not causal and (drop_rate == 0 or drop_rate == 0.0), "causal / dropout stay on the reference path")
    b, s, a, d = q.shape
    if mode == "torch":
        _require(attn_mask is None, "attn_mask stays on the reference path")
        out = attention_blhd(q, k, v)
        return out.reshape(b, s, a * d)
    # flash: q is [b, s, a, d] flattened to [(b s), a, d] by the reference's pre_attn_layout (attenion.py:22-25)
    _require(cu_seqlens_q is not None and cu_seqlens_kv is not None, "mode='flash' needs cu_seqlens")
    s1 = k.shape[1]
    qp, kp, vp = q.reshape(1, b * s, a, d), k.reshape(1, b * s1, a, d), v.reshape(1, b * s1, a, d)
    out = attention_blhd(qp, kp, vp, cu_seqlens_q=cu_seqlens_q, cu_seqlens_k=cu_seqlens_kv,
                         max_seqlen_q=max_seqlen_q, max_seqlen_k=max_seqlen_kv)
    return out.view(batch_size, max_seqlen_q, a, d).reshape(batch_size, max_seqlen_q, a * d)


# ---------------------------------------------------------------------------------------------------------------------
# Wan2.1: flash_attention(...)  (videotuna/models/wan/wan/modules/attention.py:24-130)
# ---------------------------------------------------------------------------------------------------------------------
def wan_flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None,
                        causal=False, window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
    """Same signature as the reference. q [B,Lq,N,C], k/v [B,Lk,N,C]; non-half inputs are cast to `dtype`
    (attention.py:59-83); the result comes back in q's original dtype (attention.py:57,130)."""
    _require(dtype == torch.bfloat16, "only bfloat16 compute is on the CUDA path")
    _require(not causal and tuple(window_size) == (-1, -1) and dropout_p == 0, "causal/window/dropout: reference path")
    _require(q.is_cuda, "CPU tensors stay on the reference path")
    out_dtype = q.dtype

    def half(x):
        return x if x.dtype == torch.bfloat16 else x.to(dtype)

    qh, kh, vh = half(q), half(k), half(v)
    _require(qh.shape[-1] in (64, 128), f"head dim {qh.shape[-1]} stays on the reference path")
    scale = 1.0 / math.sqrt(qh.shape[-1]) if softmax_scale is None else float(softmax_scale)
    if q_scale is not None:
        scale = scale * float(q_scale)
    out = attention_blhd(qh, kh, vh, softmax_scale=scale, k_lens=k_lens)
    return out.type(out_dtype)


# ---------------------------------------------------------------------------------------------------------------------
# lvdm: CrossAttention.forward  (videotuna/models/lvdm/modules/attention.py:101-170)
# ---------------------------------------------------------------------------------------------------------------------
def _lora_parts(layer):
    """(weight, bias, [(A, B, scaling), ...]) of a projection when it can be evaluated as base GEMM + dense-delta GEMM
    (lora_merged_projections), else None.

    A plain nn.Linear has no adapters. A peft `lora.Linear` (what `peft.get_peft_model(self.model, LoraConfig(r=4, lora_alpha=1,
    target_modules=["to_k", "to_v", "to_q"], lora_dropout=0.0))` makes of the projections, lvdm/ddpm3d.py:112-117, 436-440;
    configs/001_videocrafter2/vc2_t2v_lora.yaml:7-12) qualifies when its forward is exactly
    `base(x) + sum_a lora_B[a](lora_A[a](x)) * scaling[a]`: adapters enabled and not merged, dropout 0 (nn.Identity),
    no DoRA, no adapter bias. Anything else returns None and the caller runs the module itself."""
    if type(layer) is torch.nn.Linear:
        return layer.weight, layer.bias, []
    base = getattr(layer, "base_layer", None)
    A, B = getattr(layer, "lora_A", None), getattr(layer, "lora_B", None)
    if type(base) is not torch.nn.Linear or A is None or B is None:
        return None
    if getattr(layer, "merged", False) or getattr(layer, "disable_adapters", False):
        return None
    if isinstance(A, torch.nn.ModuleDict):
        names = list(getattr(layer, "active_adapters", A.keys()))
        if any(n not in A for n in names):
            return None
        drop, dora = getattr(layer, "lora_dropout", {}), getattr(layer, "use_dora", {})
        parts = []
        for n in names:
            if n in drop and not isinstance(drop[n], torch.nn.Identity):
                return None
            if (dora.get(n, False) if isinstance(dora, dict) else dora) or getattr(B[n], "bias", None) is not None:
                return None
            parts.append((A[n].weight, B[n].weight, float(layer.scaling[n])))
        return base.weight, base.bias, parts
    if getattr(A, "bias", None) is not None or getattr(B, "bias", None) is not None:
        return None
    return base.weight, base.bias, [(A.weight, B.weight, float(layer.scaling))]


class _LoRADeltaLinear(torch.autograd.Function):
    """y = x W^T + x D^T (+ bias) with W frozen and D = sum_a s_a B_a A_a the DENSE adapter delta (out x in):
    two GEMMs, the second accumulating into the first's output in its epilogue (beta = 1) — W and D are rounded to bf16
    separately, like the reference rounds the base weight and the adapter factors separately, so a delta far below one
    bf16 ulp of W still reaches the output. Backward: dx = dy W + dy D (same pair), dD = dy^T x (one GEMM, fp32 out)."""

    @staticmethod
    def forward(ctx, x, w, delta, bias):
        x2 = x.reshape(-1, x.shape[-1])
        y = torch.addmm(bias, x2, w.t()) if bias is not None else torch.mm(x2, w.t())
        y.addmm_(x2, delta.t())
        ctx.save_for_backward(x2, w, delta)
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w, delta = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.mm(dy2, w)
            dx.addmm_(dy2, delta)
            dx = dx.view(ctx.x_shape)
        try:
            dd = torch.mm(dy2.t(), x2, out_dtype=torch.float32)
        except (TypeError, RuntimeError):
            dd = torch.mm(dy2.t(), x2)
        return dx, None, dd.to(delta.dtype), None


def lora_merged_projections(x: Tensor, layers) -> Optional[tuple]:
    """Several LoRA projections of ONE input (to_q / to_k / to_v of self-attention; to_k / to_v of cross-attention) as dense
    GEMMs:   [y_1 | y_2 | ...] = x [W_1; W_2; ...]^T + x [D_1; D_2; ...]^T,   D_i = s_i B_i A_i   (out x in: tiny next to the
    activations). Same function and gradients as `base(x) + lora_B(lora_A(x)) * s` — dA = s B^T (dy^T x), dB = s (dy^T x) A^T
    by autograd through D — but the rank-4 GEMMs over the 81 920-row activations (three per projection and direction,
    memory-bound and run by cuBLAS at a few % of either roofline: 14 % of the VideoCrafter2 LoRA step,
    profiles/r2_s31_vc2_profile_cl_nockpt.txt), their scale / add kernels and the three-way input-gradient sum are gone.
    Returns the outputs as views of one (…, sum out) tensor, or None when a layer does not qualify (_lora_parts)."""
    parts = [_lora_parts(l) for l in layers]
    if any(p_ is None for p_ in parts) or not any(p_[2] for p_ in parts):
        return None
    if any((p_[1] is None) != (parts[0][1] is None) for p_ in parts):
        return None
    if any(p_[0].requires_grad or (p_[1] is not None and p_[1].requires_grad) for p_ in parts):
        return None  # trainable base weights: the modules' own path
    if any(p_[0].shape[0] * p_[0].shape[1] > 1024 * (p_[0].shape[0] + p_[0].shape[1]) for p_ in parts):
        # The dense delta costs 2 * rows * in * out extra FLOPs per GEMM where the rank-r path moves ~rows * (in + out)
        # elements: a win for lvdm's 320..1280-wide projections, a loss for e.g. HunyuanVideo's 3072 -> 9216 qkv.
        return None
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    deltas = []
    for w, _b, adapters in parts:
        d = None
        for a, b, s in adapters:
            t = torch.mm(b.float(), a.float()) * s
            d = t if d is None else d + t
        deltas.append(d if d is not None else torch.zeros(w.shape, device=w.device, dtype=torch.float32))
    sizes = [p_[0].shape[0] for p_ in parts]
    many = len(parts) > 1
    w = (torch.cat([p_[0] for p_ in parts], 0) if many else parts[0][0]).to(dt)
    delta = (torch.cat(deltas, 0) if many else deltas[0]).to(dt)
    bias = None if parts[0][1] is None else (torch.cat([p_[1] for p_ in parts], 0) if many else parts[0][1]).to(dt)
    y = _LoRADeltaLinear.apply(x.to(dt), w, delta, bias)
    return y.split(sizes, dim=-1) if many else (y,)


def _lvdm_qkv(self, x, context, is_self_attn):
    """to_q(x), to_k(context), to_v(context): merged-LoRA GEMMs where the projections are adapters (B200VT_LORA_MERGE=0: the
    modules' own forwards), the modules themselves otherwise."""
    if _LORA_MERGE:
        if is_self_attn:
            qkv = lora_merged_projections(x, (self.to_q, self.to_k, self.to_v))
            if qkv is not None:
                return qkv
        else:
            q, kv = lora_merged_projections(x, (self.to_q,)), lora_merged_projections(context, (self.to_k, self.to_v))
            if q is not None and kv is not None:
                return q[0], kv[0], kv[1]
    return self.to_q(x), self.to_k(context), self.to_v(context)


def lvdm_cross_attention_forward(self, x, context=None, mask=None):
    """Drop-in body for lvdm `CrossAttention.forward(self, x, context=None, mask=None)`.

    Projections stay the module's own nn.Linear layers (state-dict keys and peft LoRA targets are untouched); the
    einsum/softmax/einsum core (attention.py:126-149) runs in one CUDA kernel on the (B, N, H, D) view of the projected
    tensors, so the reference's two `rearrange` copies disappear. Sequences of at most 32 tokens (the temporal
    transformer's t = 16 frames) go to the one-warp-per-sequence kernel, which also takes the causal mask.
    """
    is_self_attn = context is None
    h = self.heads
    context = x if context is None else context
    k_ip = v_ip = None
    if self.img_cross_attention and not is_self_attn:
        context, context_img = context[:, : self.text_context_len, :], context[:, self.text_context_len:, :]
        q, k, v = _lvdm_qkv(self, x, context, False)
        k_ip, v_ip = self.to_k_ip(context_img), self.to_v_ip(context_img)
    else:
        if not is_self_attn:
            context = context[:, : self.text_context_len, :]
        q, k, v = _lvdm_qkv(self, x, context, is_self_attn)
    _require(q.is_cuda and q.dtype == torch.bfloat16, "fp32 / CPU activations stay on the reference path")
    _require(self.dim_head in (64, 128), f"dim_head {self.dim_head} stays on the reference path")
    b, n, _ = q.shape
    d = self.dim_head
    q4, k4, v4 = q.view(b, n, h, d), k.view(b, k.shape[1], h, d), v.view(b, v.shape[1], h, d)
    small = n <= 32 and k4.shape[1] == n  # the one-warp kernel takes a single N for q, k and v
    if self.relative_position:
        _require(small and d + n <= 128 and k_ip is None,
                 "relative-position attention is on the CUDA path for temporal self-attention with dim_head + frames <= 128")
        m2 = None
        if mask is not None:
            m2 = mask if mask.dim() == 2 else mask[0]
            if mask.dim() == 3 and mask.shape[0] > 1:
                _require(bool((mask == mask[:1]).all()), "per-sample masks stay on the reference path")
        out = _relative_position_attention(self, q4, k4, v4, m2).reshape(b, n, h * d)
        return self.to_out(out)
    if mask is not None:
        # TemporalTransformer's causal mask: one (t, t) pattern repeated over the batch (attention.py:487-489)
        _require(small, "masked attention is only on the CUDA path for the temporal (N <= 32) kernel")
        m = mask if mask.dim() == 2 else mask[0]
        if mask.dim() == 3 and mask.shape[0] > 1:
            _require(bool((mask == mask[:1]).all()), "per-sample masks stay on the reference path")
        out = ops.temporal_attn_fwd(q4, k4, v4, m, float(self.scale))
    elif small:
        out = ops.temporal_attn_fwd(q4, k4, v4, None, float(self.scale))
    else:
        out = attention_blhd(q4, k4, v4, softmax_scale=self.scale)
    out = out.reshape(b, n, h * d)
    if k_ip is not None:
        out_ip = attention_blhd(q4, k_ip.view(b, -1, h, d), v_ip.view(b, -1, h, d), softmax_scale=self.scale)
        out_ip = out_ip.reshape(b, n, h * d)
        if self.img_cross_attention_scale_learnable:
            out = out + self.img_cross_attention_scale * out_ip * (torch.tanh(self.alpha) + 1)
        else:
            out = out + self.img_cross_attention_scale * out_ip
    return self.to_out(out)


def _relative_position_attention(self, q4: Tensor, k4: Tensor, v4: Tensor, mask: Optional[Tensor]) -> Tensor:
    """lvdm CrossAttention with relative_position=True (VideoCrafter1, attention.py:19-42, 129-133, 145-148):
        S_ij = scale * (q_i . k_j + q_i . k2_ij),   O_i = sum_j P_ij (v_j + v2_ij),   k2_ij = table_k[j - i + max], v2 likewise.
    Both extra terms ride on the one-warp temporal kernel at head dim 128 through an AUGMENTED head: with e_j the one-hot
    of key position j,
        q'_i = [q_i ; m_i ; 0],  m_i[j] = q_i . k2_ij        k'_j = [k_j ; e_j ; 0]   =>  q'_i . k'_j = q_i . k_j + q_i . k2_ij
        v'_j = [v_j ; e_j ; 0]                               =>  O'_i = [ sum_j P_ij v_j ; P_i. ; 0 ]
    so the kernel's extra output columns ARE the attention probabilities, from which sum_j P_ij v2_ij is a small einsum.
    Autograd flows through the same kernel backward (the e_j columns are constants). q4, k4, v4: (B, N, H, D), D + N <= 128."""
    b, n, h, d = q4.shape
    k2 = self.relative_position_k(n, n).to(q4.dtype)   # (N, N, D); the module call keeps the table's gradient
    v2 = self.relative_position_v(n, n).to(q4.dtype)
    m = torch.einsum("bthd,tsd->bths", q4, k2)         # (B, N, H, N)
    eye = torch.eye(n, device=q4.device, dtype=q4.dtype).view(1, n, 1, n).expand(b, n, h, n)
    pad = q4.new_zeros((b, n, h, 128 - d - n))
    q_aug = torch.cat([q4, m, pad], dim=-1)
    k_aug = torch.cat([k4, eye, pad], dim=-1)
    v_aug = torch.cat([v4, eye, pad], dim=-1)
    o_aug = ops.temporal_attn_fwd(q_aug, k_aug, v_aug, mask, float(self.scale))
    return o_aug[..., :d] + torch.einsum("bths,tsd->bthd", o_aug[..., d:d + n], v2)


def temporal_attention(q: Tensor, k: Tensor, v: Tensor, softmax_scale: Optional[float] = None,
                       mask: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T * scale [mask]) v for (B, N, H, D) tensors with N <= 32: lvdm TemporalTransformer's attention
    over frames (attention.py:475-519). mask: (N, N), > 0.5 = keep."""
    _require(supported_qkv(q, k, v) and q.shape[1] <= 32 and k.shape[1] <= 32 and q.shape[1] == k.shape[1],
             "temporal attention needs CUDA bf16 (B, N<=32, H, D in {64,128}) tensors")
    scale = 1.0 / math.sqrt(q.shape[-1]) if softmax_scale is None else float(softmax_scale)
    return ops.temporal_attn_fwd(q, k, v, mask, scale)


# ---------------------------------------------------------------------------------------------------------------------
# memory-bound helpers
# ---------------------------------------------------------------------------------------------------------------------
def ln_modulate(x: Tensor, shift: Optional[Tensor] = None, scale: Optional[Tensor] = None,
                weight: Optional[Tensor] = None, bias: Optional[Tensor] = None, eps: float = 1e-6,
                tr_shift: Optional[Tensor] = None, tr_scale: Optional[Tensor] = None,
                first_frame_tokens: int = 0) -> Tensor:
    """modulate(LayerNorm(x), shift, scale) in one pass (hunyuan modulate_layers.py:31-49 after nn.LayerNorm;
    wan model.py:294-296). x (B,L,C) bf16 or fp32 (Wan's fp32 residual stream); the result is bf16, ready for the
    following Linear; shift/scale (B,C) or (B,1,C).
    first_frame_tokens > 0: the i2v "token_replace" form (hyvideo_i2v/modules/modulate_layers.py:37-63) — rows
    [0, first_frame_tokens) of every sample take tr_shift / tr_scale, the rest shift / scale."""
    _require(x.is_cuda and x.dtype in (torch.bfloat16, torch.float32) and x.dim() == 3 and x.shape[-1] % 8 == 0,
             "ln_modulate needs a CUDA bf16 or fp32 (B,L,C) tensor")
    B, _, Cc = x.shape
    sc = None if scale is None else scale.reshape(B, Cc)
    sh = None if shift is None else shift.reshape(B, Cc)
    if first_frame_tokens > 0:
        _require(tr_shift is not None and tr_scale is not None and sc is not None and sh is not None,
                 "token_replace modulation needs both vector sets")
        y, _, _ = ops.ln_modulate_fwd(x, weight, bias, sc, sh, float(eps), tr_scale.reshape(B, Cc),
                                      tr_shift.reshape(B, Cc), int(first_frame_tokens))
        return y
    y, _, _ = ops.ln_modulate_fwd(x, weight, bias, sc, sh, float(eps))
    return y


def layer_norm(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], eps: float) -> Tensor:
    """nn.LayerNorm over the last dim for (.., C) bf16 tensors (lvdm BasicTransformerBlock norms, attention.py:262-264)."""
    shp = x.shape
    y = ln_modulate(x.reshape(1, -1, shp[-1]), None, None, weight, bias, eps)
    return y.view(shp)


def gate_residual(x: Tensor, branch: Tensor, gate: Optional[Tensor] = None, tr_gate: Optional[Tensor] = None,
                  first_frame_tokens: int = 0) -> Tensor:
    """x + apply_gate(branch, gate) (hunyuan modulate_layers.py:52-68, models.py:231; wan model.py:298).
    first_frame_tokens > 0: i2v "token_replace" (hyvideo_i2v/modules/modulate_layers.py:66-96) — rows
    [0, first_frame_tokens) of every sample are gated by tr_gate."""
    _require(x.is_cuda and x.dtype in (torch.bfloat16, torch.float32) and branch.dtype == torch.bfloat16
             and x.dim() == 3, "gate_residual needs a CUDA bf16/fp32 (B,L,C) x and a bf16 branch")
    B, _, Cc = x.shape
    g = None if gate is None else gate.reshape(B, Cc)
    if first_frame_tokens > 0 and g is not None:
        _require(tr_gate is not None, "token_replace gating needs tr_gate")
        return ops.gate_residual_fwd(x, branch, g, tr_gate.reshape(B, Cc), int(first_frame_tokens))
    return ops.gate_residual_fwd(x, branch, g)


def qk_rmsnorm_rope(x: Tensor, weight: Optional[Tensor], cos: Optional[Tensor], sin: Optional[Tensor],
                    per_head: bool = True, eps: float = 1e-6) -> Tensor:
    """RMSNorm (per head, or over the whole token when per_head=False) followed by interleaved RoPE on the first
    cos.shape[0] tokens. x (B,L,H,D) bf16 (strided views allowed); cos/sin (L_rope, D) fp32 (posemb_layers.py:140-171,
    norm_layers.py:33-59; wan model.py:40-86)."""
    _require(x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[-1] in (64, 128),
             "qk_rmsnorm_rope needs a CUDA bf16 (B,L,H,D) tensor with D in {64,128}")
    mode = 0 if weight is None else (1 if per_head else 2)
    y, _ = ops.qk_rmsnorm_rope_fwd(x, weight, cos, sin, mode, float(eps))
    return y


def lvdm_geglu_forward(self, x: Tensor) -> Tensor:
    """Drop-in body of lvdm GEGLU.forward (attention.py:527-529): the module's own projection, then x * gelu(gate) in one
    pass over its output (the widest tensor of the UNet) instead of chunk / gelu / mul."""
    h = self.proj(x)
    _require(h.is_cuda and h.dtype == torch.bfloat16 and h.shape[-1] % 16 == 0,
             "fused GEGLU needs a CUDA bf16 projection output (bf16 weights or autocast)")
    return ops.geglu_fwd(h)


def hunyuan_joint_qkv(img_qkv: Tensor, txt_qkv: Tensor, img_q_norm, img_k_norm, txt_q_norm, txt_k_norm,
                      cos: Optional[Tensor], sin: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """MMDoubleStreamBlock's q/k/v preparation (hyvideo_t2v/modules/models.py:166-194) in one op: img_qkv (B,L,3,H,D) and
    txt_qkv (B,T,3,H,D) -> joint q, k, v (B, L+T, H, D) with the per-head RMSNorm of both streams (norm_layers.py:33-59)
    and the RoPE of the image rows (posemb_layers.py:140-188) applied on the way; replaces the three torch.cat.
    *_norm: the block's RMSNorm modules or nn.Identity (qk_norm=False)."""
    _require(img_qkv.is_cuda and img_qkv.dtype == torch.bfloat16 and txt_qkv.dtype == torch.bfloat16
             and img_qkv.shape[-1] in (64, 128), "joint q/k/v needs CUDA bf16 tensors with head dim 64/128")
    ws, eps = [], None
    for n in (img_q_norm, img_k_norm, txt_q_norm, txt_k_norm):
        if isinstance(n, torch.nn.Identity):
            ws.append(None)
            continue
        _require(hasattr(n, "weight") and hasattr(n, "eps") and not isinstance(n, torch.nn.LayerNorm),
                 "only RMSNorm q/k normalisation is on the CUDA path")
        _require(eps is None or float(n.eps) == eps, "q/k norms with different eps stay on the reference path")
        eps = float(n.eps)
        ws.append(n.weight)
    _require((ws[0] is None) == (ws[1] is None) and (ws[2] is None) == (ws[3] is None), "q and k norms must match")
    q, k, v, _, _ = ops.joint_qkv_fwd(img_qkv, txt_qkv, ws[0], ws[1], ws[2], ws[3], cos, sin, 1e-6 if eps is None else eps)
    return q, k, v


def groupnorm_silu(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], groups: int, eps: float,
                   silu: bool = False, addend: Optional[Tensor] = None) -> Tensor:
    """GroupNorm with fp32 statistics (+ SiLU) on (N,C,*) bf16/fp32 tensors (lvdm GroupNormSpecific, utils.py:192-203).
    addend: optional fp32 (C,) or (N, C) term, y = GroupNorm(x + addend[..., None, None]) — a convolution bias and / or
    ResBlock's timestep embedding folded into the kernel's per-channel constants (channels-last x only; no gradient flows
    to it: pass it only when it needs none)."""
    _require(x.is_cuda and x.dtype in (torch.bfloat16, torch.float32), "groupnorm needs a CUDA bf16/fp32 tensor")
    if addend is not None:
        _require(not (addend.requires_grad and torch.is_grad_enabled()), "a GroupNorm addend that needs a gradient is added by the caller")
        y, _, _ = ops.groupnorm_silu_fwd(x, weight, bias, int(groups), float(eps), bool(silu), addend.detach().float())
        return y
    y, _, _ = ops.groupnorm_silu_fwd(x, weight, bias, int(groups), float(eps), bool(silu), None)
    return y
