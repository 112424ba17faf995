"""Denoiser-level iterations/s (BASELINE.json metric, second half: "DiT train it/s at 1/2/4/8 B200"; SURVEY.md §8(d)
"End-to-end"): forward + backward of the transformer-block stack of a DiT backbone on synthetic tokens of the named
shape, with random-init weights of the named architecture.

    python tools/bench_denoiser.py --model hunyuan|wan|cogvideox [--arm ours|torch] [--double N --single M | --layers N]
                                   [--steps K] [--warmup W] [--no-checkpoint]
    python -m torch.distributed.run --nproc-per-node P ... tools/bench_denoiser.py --model hunyuan   (Ulysses SP)

arms
  ours    the reference blocks' constructors (tests/helpers.py shells: same sub-module and parameter names as
          MMDoubleStreamBlock / MMSingleStreamBlock / WanAttentionBlock) with the forwards patch.patch_blocks() installs
          (b200vt.blocks): fused LayerNorm+modulate, QK-RMSNorm+RoPE, tcgen05 attention, gated residual.
  torch   the same modules and weights driven by the reference's own sequence of torch ops, written out here
          (hunyuan models.py:132-252, 326-393 with attention mode="flash" -> flash_attn_varlen_func, attenion.py:108-123;
          wan model.py:127-156, 274-313 with flash_attention -> flash_attn_varlen_func, attention.py:113-130), i.e. what
          the unpatched reference launches on this GPU. Single GPU only.
Both arms: bf16 weights, per-block activation checkpointing (use_reentrant=False, as lvdm/utils.py:122 and the i2v DiT
switch hyvideo_i2v/modules/models.py:764-769 do), hunyuan = LoRA step (frozen base weights, rank-4 adapters on the
attention projections, configs/007_hunyuanvideo/hunyuanvideo_t2v_diffuser_lora.yaml:56-61), wan = full fwd+bwd (weight
gradients on), cogvideox = full fwd+bwd of diffusers-style CogVideoXBlocks (torch arm: the stock block math with
F.scaled_dot_product_attention, as CogVideoXAttnProcessor2_0 calls it). Patch embedding, text refiner and final layer (< 0.1 % of the FLOPs) are not part of the stack: the inputs
are the token streams the first block sees. One iteration = forward + backward of the stack; no optimizer step (rank-4
adapters / out of the hot path). Timing: CUDA events around K iterations, barrier + synchronize on both sides, max over
ranks. This is a tool next to bench.py (whose contract stays the attention metric), not a replacement for it."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn
from torch.utils.checkpoint import checkpoint

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vt._lib as L  # noqa: E402
import b200vt.sp as sp  # noqa: E402
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("b200vt_test_helpers", os.path.join(ROOT, "tests", "helpers.py"))
H = importlib.util.module_from_spec(_spec)  # shells of the reference constructors
_spec.loader.exec_module(H)

SEED = 20230211
BF16 = torch.bfloat16

CONFIGS = {
    # HunyuanVideo T2V 720x1280x129 (C4): hyvideo_t2v/modules/models.py:742-750 (HUNYUAN_VIDEO_CONFIG "HYVideo-T/2-cfgdistill")
    "hunyuan": dict(hidden=3072, heads=24, mlp_ratio=4.0, double=20, single=40, grid=(33, 45, 80), txt=256,
                    rope_dims=(16, 56, 56), theta=256.0),
    # Wan2.1-T2V-14B 480x832x81 (C5): wan/configs/wan_t2v_14B.py:20-29
    "wan": dict(dim=5120, ffn=13824, heads=40, layers=40, grid=(21, 30, 52), txt=512),
    # CogVideoX-2B 480x720x49 (C3): diffusers 0.32.2 CogVideoXTransformer3DModel config of THUDM/CogVideoX-2b [ext]
    # (30 heads x 64, 30 layers, time_embed_dim 512, text 226 tokens, no rotary embedding in the 2B model)
    "cogvideox": dict(dim=1920, heads=30, layers=30, time_embed_dim=512, grid=(13, 30, 45), txt=226),
}


# =====================================================================================================================
# LoRA (peft 0.12 semantics: y = base(x) + B(A(x)) * alpha / r; lora_B re-drawn N(0, 0.02) so its gradient path is live)
# =====================================================================================================================
class LoRALinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int = 4, alpha: float = 1.0):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.Linear(base.in_features, r, bias=False, device=base.weight.device, dtype=base.weight.dtype)
        self.lora_B = nn.Linear(r, base.out_features, bias=False, device=base.weight.device, dtype=base.weight.dtype)
        nn.init.normal_(self.lora_B.weight, std=0.02)
        self.scaling = alpha / r

    def forward(self, x):
        return self.base_layer(x) + self.lora_B(self.lora_A(x)) * self.scaling


def add_lora(block: nn.Module, names) -> None:
    for p in block.parameters():
        p.requires_grad_(False)
    for n in names:
        setattr(block, n, LoRALinear(getattr(block, n)))


# =====================================================================================================================
# torch arm: the reference's op sequence
# =====================================================================================================================
def _flash_varlen(q, k, v, cu_q, cu_k, max_q, max_k):
    from flash_attn import flash_attn_varlen_func
    return flash_attn_varlen_func(q, k, v, cu_q, cu_k, max_q, max_k)


def hy_modulate(x, shift, scale):  # modulate_layers.py:31-49
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def hy_rmsnorm(norm, x):  # norm_layers.py:33-59
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + norm.eps)).type_as(x) * norm.weight


def hy_rope(x, cos, sin):  # posemb_layers.py:140-188, real-valued branch
    xf = x.float()
    a, b = xf.reshape(*xf.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-b, a], dim=-1).flatten(3)
    return (xf * cos.view(1, -1, 1, cos.shape[-1]) + rot * sin.view(1, -1, 1, sin.shape[-1])).type_as(x)


def hy_attention_flash(q, k, v, cu_q, cu_k, max_q, max_k, batch):  # attenion.py:60-156, mode="flash"
    b, s, h, d = q.shape
    x = _flash_varlen(q.reshape(b * s, h, d), k.reshape(b * k.shape[1], h, d), v.reshape(b * v.shape[1], h, d),
                      cu_q, cu_k, max_q, max_k)
    return x.view(batch, max_q, h, d).reshape(batch, max_q, h * d)


def hy_double_torch(m, img, txt, vec, cu_q, cu_k, max_q, max_k, freqs_cis):  # models.py:132-252
    B, Lq, C = img.shape
    Hh = m.heads_num
    i_sh1, i_sc1, i_g1, i_sh2, i_sc2, i_g2 = m.img_mod(vec).chunk(6, dim=-1)
    t_sh1, t_sc1, t_g1, t_sh2, t_sc2, t_g2 = m.txt_mod(vec).chunk(6, dim=-1)
    img_qkv = m.img_attn_qkv(hy_modulate(m.img_norm1(img), i_sh1, i_sc1))
    img_q, img_k, img_v = img_qkv.view(B, Lq, 3, Hh, -1).permute(2, 0, 1, 3, 4).unbind(0)
    img_q, img_k = hy_rmsnorm(m.img_attn_q_norm, img_q).to(img_v), hy_rmsnorm(m.img_attn_k_norm, img_k).to(img_v)
    img_q, img_k = hy_rope(img_q, *freqs_cis), hy_rope(img_k, *freqs_cis)
    txt_qkv = m.txt_attn_qkv(hy_modulate(m.txt_norm1(txt), t_sh1, t_sc1))
    txt_q, txt_k, txt_v = txt_qkv.view(B, txt.shape[1], 3, Hh, -1).permute(2, 0, 1, 3, 4).unbind(0)
    txt_q, txt_k = hy_rmsnorm(m.txt_attn_q_norm, txt_q).to(txt_v), hy_rmsnorm(m.txt_attn_k_norm, txt_k).to(txt_v)
    q, k, v = torch.cat((img_q, txt_q), 1), torch.cat((img_k, txt_k), 1), torch.cat((img_v, txt_v), 1)
    attn = hy_attention_flash(q, k, v, cu_q, cu_k, max_q, max_k, B)
    img_attn, txt_attn = attn[:, :Lq], attn[:, Lq:]
    img = img + m.img_attn_proj(img_attn) * i_g1.unsqueeze(1)
    img = img + m.img_mlp(hy_modulate(m.img_norm2(img), i_sh2, i_sc2)) * i_g2.unsqueeze(1)
    txt = txt + m.txt_attn_proj(txt_attn) * t_g1.unsqueeze(1)
    txt = txt + m.txt_mlp(hy_modulate(m.txt_norm2(txt), t_sh2, t_sc2)) * t_g2.unsqueeze(1)
    return img, txt


def hy_single_torch(m, x, vec, txt_len, cu_q, cu_k, max_q, max_k, freqs_cis):  # models.py:326-393
    B, S, C = x.shape
    Hh = m.heads_num
    sh, sc, gate = m.modulation(vec).chunk(3, dim=-1)
    qkv, mlp = torch.split(m.linear1(hy_modulate(m.pre_norm(x), sh, sc)), [3 * C, m.mlp_hidden_dim], dim=-1)
    q, k, v = qkv.view(B, S, 3, Hh, -1).permute(2, 0, 1, 3, 4).unbind(0)
    q, k = hy_rmsnorm(m.q_norm, q).to(v), hy_rmsnorm(m.k_norm, k).to(v)
    img_q, txt_q = q[:, :-txt_len], q[:, -txt_len:]
    img_k, txt_k = k[:, :-txt_len], k[:, -txt_len:]
    q = torch.cat((hy_rope(img_q, *freqs_cis), txt_q), dim=1)
    k = torch.cat((hy_rope(img_k, *freqs_cis), txt_k), dim=1)
    attn = hy_attention_flash(q, k, v, cu_q, cu_k, max_q, max_k, B)
    return x + m.linear2(torch.cat((attn, m.mlp_act(mlp)), 2)) * gate.unsqueeze(1)


def wan_rmsnorm(norm, x):  # model.py:70-86
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + norm.eps)).type_as(x) * norm.weight


def wan_ln(norm, x):  # model.py:89-99: fp32 LayerNorm, result in x's dtype
    w = getattr(norm, "weight", None)
    b = getattr(norm, "bias", None)
    return F.layer_norm(x.float(), (x.shape[-1],), None if w is None else w.float(), None if b is None else b.float(),
                        norm.eps).type_as(x)


def wan_rope(x, grid_sizes, freqs):  # model.py:40-67 (float64 complex multiply, fp32 result)
    n, c = x.size(2), x.size(3) // 2
    fs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    out = []
    for i, (f, h, w) in enumerate(grid_sizes.tolist()):
        s = f * h * w
        xi = torch.view_as_complex(x[i, :s].to(torch.float64).reshape(s, n, -1, 2))
        fr = torch.cat([fs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), fs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                        fs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(s, 1, -1)
        xi = torch.view_as_real(xi * fr).flatten(2)
        out.append(torch.cat([xi, x[i, s:]]))
    return torch.stack(out).float()


def wan_flash(q, k, v, k_lens=None):  # attention.py:24-130 (batch of equal lengths; output in q's dtype)
    b, lq, lk, out_dtype = q.size(0), q.size(1), k.size(1), q.dtype
    qh, kh, vh = (t.to(BF16).flatten(0, 1) for t in (q, k, v))
    cu_q = torch.arange(0, (b + 1) * lq, lq, dtype=torch.int32, device=q.device)
    cu_k = torch.arange(0, (b + 1) * lk, lk, dtype=torch.int32, device=q.device)
    return _flash_varlen(qh, kh, vh, cu_q, cu_k, lq, lk).unflatten(0, (b, lq)).type(out_dtype)


def wan_self_torch(m, x, seq_lens, grid_sizes, freqs):  # model.py:127-156
    b, s, n, d = *x.shape[:2], m.num_heads, m.head_dim
    q = wan_rmsnorm(m.norm_q, m.q(x)).view(b, s, n, d)
    k = wan_rmsnorm(m.norm_k, m.k(x)).view(b, s, n, d)
    v = m.v(x).view(b, s, n, d)
    return m.o(wan_flash(wan_rope(q, grid_sizes, freqs), wan_rope(k, grid_sizes, freqs), v).flatten(2))


def wan_cross_torch(m, x, context, context_lens):  # model.py:161-181
    b, n, d = x.size(0), m.num_heads, m.head_dim
    q = wan_rmsnorm(m.norm_q, m.q(x)).view(b, -1, n, d)
    k = wan_rmsnorm(m.norm_k, m.k(context)).view(b, -1, n, d)
    v = m.v(context).view(b, -1, n, d)
    return m.o(wan_flash(q, k, v).flatten(2))


def wan_block_torch(m, x, e, seq_lens, grid_sizes, freqs, context, context_lens):  # model.py:274-313
    e = (m.modulation.float() + e).chunk(6, dim=1)
    y = wan_self_torch(m.self_attn, wan_ln(m.norm1, x).float() * (1 + e[1]) + e[0], seq_lens, grid_sizes, freqs)
    x = x + y * e[2]
    x = x + wan_cross_torch(m.cross_attn, wan_ln(m.norm3, x), context, context_lens)
    y = m.ffn(wan_ln(m.norm2, x).float() * (1 + e[4]) + e[3])
    return x + y * e[5]


def cog_norm_zero_torch(norm, h, e, temb):  # diffusers CogVideoXLayerNormZero.forward [ext]
    shift, scale, gate, e_shift, e_scale, e_gate = norm.linear(norm.silu(temb)).chunk(6, dim=1)
    h = norm.norm(h) * (1 + scale)[:, None, :] + shift[:, None, :]
    e = norm.norm(e) * (1 + e_scale)[:, None, :] + e_shift[:, None, :]
    return h, e, gate[:, None, :], e_gate[:, None, :]


def cog_attn_torch(attn, h, e):  # diffusers CogVideoXAttnProcessor2_0.__call__ [ext]: SDPA on (B, heads, S, d)
    T = e.size(1)
    x = torch.cat([e, h], dim=1)
    B, S, _ = x.shape
    q, k, v = (lin(x).view(B, S, attn.heads, -1).transpose(1, 2) for lin in (attn.to_q, attn.to_k, attn.to_v))
    q, k = attn.norm_q(q), attn.norm_k(k)
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
    o = attn.to_out[1](attn.to_out[0](o.transpose(1, 2).reshape(B, S, -1)))
    return o[:, T:], o[:, :T]


def cog_block_torch(m, h, e, temb):  # diffusers CogVideoXBlock.forward [ext]
    T = e.size(1)
    hn, en, g, eg = cog_norm_zero_torch(m.norm1, h, e, temb)
    ah, ae = cog_attn_torch(m.attn1, hn, en)
    h = h + g * ah
    e = e + eg * ae
    hn, en, g, eg = cog_norm_zero_torch(m.norm2, h, e, temb)
    ff = m.ff(torch.cat([en, hn], dim=1))
    return h + g * ff[:, T:], e + eg * ff[:, :T]


# =====================================================================================================================
# tables
# =====================================================================================================================
def hunyuan_rope_tables(rope_dims, grid, theta, device):  # posemb_layers.py:191-310 (use_real, no interpolation)
    axes = torch.meshgrid(*[torch.arange(n, dtype=torch.float32) for n in grid], indexing="ij")
    cs, sn = [], []
    for d, pos in zip(rope_dims, axes):
        inv = 1.0 / (theta ** (torch.arange(0, d, 2)[: d // 2].float() / d))
        ang = torch.outer(pos.reshape(-1), inv)
        cs.append(ang.cos().repeat_interleave(2, dim=1))
        sn.append(ang.sin().repeat_interleave(2, dim=1))
    return torch.cat(cs, 1).to(device), torch.cat(sn, 1).to(device)


def wan_freqs_table(d, device):  # model.py:29-36, 469-474
    def params(dim):
        ang = torch.outer(torch.arange(1024), 1.0 / torch.pow(10000.0, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
        return torch.polar(torch.ones_like(ang), ang)
    return torch.cat([params(d - 4 * (d // 6)), params(2 * (d // 6)), params(2 * (d // 6))], dim=1).to(device)


# =====================================================================================================================
def build_hunyuan(cfg, n_double, n_single, dev):
    torch.set_default_dtype(BF16)
    try:
        with torch.device(dev):
            dbl = [H.HunyuanDoubleShell(cfg["hidden"], cfg["heads"], cfg["mlp_ratio"]) for _ in range(n_double)]
            sgl = [H.HunyuanSingleShell(cfg["hidden"], cfg["heads"], cfg["mlp_ratio"]) for _ in range(n_single)]
    finally:
        torch.set_default_dtype(torch.float32)
    for b in dbl:
        add_lora(b, ("img_attn_qkv", "img_attn_proj", "txt_attn_qkv", "txt_attn_proj"))
    for b in sgl:
        add_lora(b, ("linear1", "linear2"))
    return dbl, sgl


def build_wan(cfg, n_layers, dev):
    torch.set_default_dtype(BF16)
    try:
        with torch.device(dev):
            blocks = [H.WanBlockShell(cfg["dim"], cfg["ffn"], cfg["heads"]) for _ in range(n_layers)]
    finally:
        torch.set_default_dtype(torch.float32)
    for b in blocks:
        nn.init.normal_(b.modulation, std=0.02)  # zero-init in the reference; re-drawn (SURVEY §4 trap 1)
    return blocks


def build_cogvideox(cfg, n_layers, dev):
    torch.set_default_dtype(BF16)
    try:
        with torch.device(dev):
            return [H.CogVideoXBlockShell(cfg["dim"], cfg["heads"], cfg["time_embed_dim"]) for _ in range(n_layers)]
    finally:
        torch.set_default_dtype(torch.float32)


def parse(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model", choices=tuple(CONFIGS), default="hunyuan")
    ap.add_argument("--arm", choices=("ours", "torch"), default="ours")
    ap.add_argument("--double", type=int, default=None)
    ap.add_argument("--single", type=int, default=None)
    ap.add_argument("--layers", type=int, default=None, help="wan: number of blocks")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-checkpoint", action="store_true")
    ap.add_argument("--keep-attention", action="store_true",
                    help="ours arm, with checkpointing: selective activation checkpointing that keeps the attention outputs (O, "
                    "LSE) instead of recomputing the attention forward in the backward pass (b200vt.ckpt)")
    ap.add_argument("--graph", action="store_true",
                    help="capture the whole iteration (forward, backward, optimizer step) into one CUDA graph and replay it "
                    "(single GPU)")
    ap.add_argument("--optimizer", choices=("none", "adamw"), default="none",
                    help="adamw: one fused AdamW step on the trainable parameters (LoRA adapters for hunyuan, every block "
                    "weight for wan / cogvideox) inside the timed iteration, lr 1e-5 as the reference configs")
    ap.add_argument("--no-grad-sync", action="store_true", help="diagnostic (N > 1): skip the all-reduce of the replicated "
                    "parameters' gradients, to size its share of the sequence-parallel step")
    ap.add_argument("--tokens-scale", type=float, default=1.0, help="debug: shrink the latent frame count")
    ap.add_argument("--check", action="store_true", help="run one forward+backward of BOTH arms on the same weights and "
                    "print the relative difference of the outputs and input gradients, then exit")
    return ap.parse_args(argv)


def run(args, manage_dist: bool = True, emit: bool = True):
    """One measurement; returns the result line (rank 0) or None. manage_dist=False: the caller (bench.py) owns the process
    group and the CUDA device; emit=False: do not print."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if manage_dist:
            dist.init_process_group("nccl", device_id=dev)
        assert args.arm == "ours", "the torch arm is single-GPU"
    L.call("vt_init", local)
    torch.manual_seed(SEED)
    cfg = dict(CONFIGS[args.model])
    f, h, w = cfg["grid"]
    f = max(1, int(round(f * args.tokens_scale)))
    n_img, n_txt = f * h * w, cfg["txt"]
    assert n_img % world == 0
    n_loc = n_img // world
    ours = args.arm == "ours"
    ckpt = not args.no_checkpoint

    use_graph = bool(getattr(args, "graph", False))
    assert not (use_graph and world > 1), "--graph is single-GPU"

    keep_attn = bool(getattr(args, "keep_attention", False)) and ckpt and ours
    if keep_attn:
        import b200vt.ckpt as CK
        import b200vt.sp  # noqa: F401  (registers the sequence-parallel attention op the policy also keeps)

    def run_block(fn, *a):
        # under graph capture the checkpoint must not save / restore the RNG state (a host operation); these blocks have no dropout
        if not ckpt:
            return fn(*a)
        if keep_attn:  # selective: the attention outputs (O, LSE) stay resident, everything else is recomputed
            return checkpoint(fn, *a, use_reentrant=False, preserve_rng_state=not use_graph, context_fn=CK.context_fn)
        return checkpoint(fn, *a, use_reentrant=False, preserve_rng_state=not use_graph)

    opt_state = {}

    def finish_step(params):
        """End of an iteration: the optimizer step of a finetuning iteration (--optimizer adamw), then drop the gradients."""
        if getattr(args, "optimizer", "none") == "adamw":
            if "opt" not in opt_state:
                opt_state["opt"] = torch.optim.AdamW([p for p in params if p.requires_grad], lr=1e-5, fused=True,
                                                     capturable=use_graph)
            opt_state["opt"].step()
        for p in params:
            p.grad = None

    if args.model == "hunyuan":
        nd = cfg["double"] if args.double is None else args.double
        ns = cfg["single"] if args.single is None else args.single
        dbl, sgl = build_hunyuan(cfg, nd, ns, dev)
        C, heads = cfg["hidden"], cfg["heads"]
        assert heads % world == 0
        cos, sin = hunyuan_rope_tables(cfg["rope_dims"], (f, h, w), cfg["theta"], dev)
        cos, sin = cos[rank * n_loc:(rank + 1) * n_loc].contiguous(), sin[rank * n_loc:(rank + 1) * n_loc].contiguous()
        g = torch.Generator(device=dev).manual_seed(SEED + rank)
        img0 = torch.randn(1, n_loc, C, device=dev, dtype=BF16, generator=g)
        g0 = torch.Generator(device=dev).manual_seed(SEED)
        txt0 = torch.randn(1, n_txt, C, device=dev, dtype=BF16, generator=g0)
        vec = torch.randn(1, C, device=dev, dtype=BF16, generator=g0)
        d_out = torch.randn(1, n_loc + n_txt, C, device=dev, dtype=BF16, generator=g) * 1e-2
        S = n_loc + n_txt
        cu = torch.tensor([0, S, S], dtype=torch.int32, device=dev)  # all text tokens valid (attenion.py:34-57)
        if world > 1:
            attn_sp = sp.UlyssesAttention()
            for b in (*dbl, *sgl):
                b.hybrid_seq_parallel_attn = attn_sp
        params = [p for b in (*dbl, *sgl) for p in b.parameters() if p.requires_grad]
        n_params = sum(p.numel() for b in (*dbl, *sgl) for p in b.parameters())
        layers_desc = {"double": nd, "single": ns}
        attn_flops = 14.0 * heads * float(n_img + n_txt) ** 2 * (C // heads) * (nd + ns)
        attn_exec = attn_flops * (18.0 / 14.0 if ckpt and not keep_attn else 1.0)

        def iteration(ours=ours):
            img, txt = img0.detach().requires_grad_(True), txt0.detach().requires_grad_(True)
            leaf = img
            for b in dbl:
                fn = b if ours else (lambda *a, _b=b: hy_double_torch(_b, *a))
                img, txt = run_block(fn, img, txt, vec, cu, cu, S, S, (cos, sin))
            x = torch.cat((img, txt), 1)
            for b in sgl:
                fn = b if ours else (lambda *a, _b=b: hy_single_torch(_b, *a))
                x = run_block(fn, x, vec, n_txt, cu, cu, S, S, (cos, sin))
            x.backward(d_out)
            finish_step(params)
            return x.detach(), leaf.grad
    elif args.model == "cogvideox":
        assert world == 1, "CogVideoX-2B has 30 heads: data parallel only (SURVEY 8e); bench it at 1 GPU"
        nl = cfg["layers"] if args.layers is None else args.layers
        blocks = build_cogvideox(cfg, nl, dev)
        C, heads = cfg["dim"], cfg["heads"]
        g = torch.Generator(device=dev).manual_seed(SEED)
        h0 = torch.randn(1, n_img, C, device=dev, dtype=BF16, generator=g)
        t0 = torch.randn(1, n_txt, C, device=dev, dtype=BF16, generator=g)
        temb = torch.randn(1, cfg["time_embed_dim"], device=dev, dtype=BF16, generator=g)
        d_out = torch.randn(1, n_img, C, device=dev, dtype=BF16, generator=g) * 1e-2
        params = [p for b in blocks for p in b.parameters()]
        n_params = sum(p.numel() for p in params)
        layers_desc = {"layers": nl}
        attn_flops = 14.0 * heads * float(n_img + n_txt) ** 2 * (C // heads) * nl
        attn_exec = attn_flops * (18.0 / 14.0 if ckpt and not keep_attn else 1.0)

        def iteration(ours=ours):
            h = leaf = h0.detach().requires_grad_(True)
            e = t0.detach().requires_grad_(True)
            for b in blocks:
                fn = b if ours else (lambda *a, _b=b: cog_block_torch(_b, *a))
                h, e = run_block(fn, h, e, temb)
            torch.autograd.backward([h, e], [d_out, torch.zeros_like(e)])
            finish_step(params)
            return h.detach(), leaf.grad
    else:
        nl = cfg["layers"] if args.layers is None else args.layers
        blocks = build_wan(cfg, nl, dev)
        C, heads = cfg["dim"], cfg["heads"]
        assert heads % world == 0
        g = torch.Generator(device=dev).manual_seed(SEED + rank)
        x0 = torch.randn(1, n_loc, C, device=dev, dtype=torch.float32, generator=g)
        d_out = torch.randn(1, n_loc, C, device=dev, dtype=torch.float32, generator=g) * 1e-2
        g0 = torch.Generator(device=dev).manual_seed(SEED)
        e0 = torch.randn(1, 6, C, device=dev, dtype=torch.float32, generator=g0) * 0.1
        ctx = torch.randn(1, n_txt, C, device=dev, dtype=BF16, generator=g0)
        seq_lens = torch.tensor([n_img], dtype=torch.long, device=dev)
        grid_sizes = torch.tensor([[f, h, w]], dtype=torch.long)
        freqs = wan_freqs_table(C // heads, dev)
        if world > 1:
            # sequence parallel as the reference wires it (wan/text2video.py:261-271): every block's self-attention forward is
            # rebound to the Ulysses version; tokens are sharded, the text context is replicated, and the weight gradients of
            # the replicated parameters are summed over the ranks after the backward (one all-reduce per parameter)
            import b200vt.patch as P
            for b in blocks:
                b.self_attn._fwd = lambda self_, *a, **k: P.wan_usp_attn_forward(self_, *a, **k)
        params = [p for b in blocks for p in b.parameters()]
        n_params = sum(p.numel() for p in params)
        layers_desc = {"layers": nl}
        attn_flops = 14.0 * heads * (float(n_img) ** 2 + float(n_img) * n_txt) * (C // heads) * nl
        attn_exec = attn_flops * (18.0 / 14.0 if ckpt and not keep_attn else 1.0)

        def iteration(ours=ours):
            x = leaf = x0.detach().requires_grad_(True)
            with torch.autocast("cuda", dtype=BF16):
                for b in blocks:
                    fn = b if ours else (lambda *a, _b=b: wan_block_torch(_b, *a))
                    x = run_block(fn, x, e0, seq_lens, grid_sizes, freqs, ctx, None)
            x.backward(d_out)
            for hnd in pending:
                hnd.wait()
            pending.clear()
            finish_step(params)
            return x.detach(), leaf.grad

        pending = []
        if world > 1:
            # DDP-style: a parameter's gradient is all-reduced as soon as it has been accumulated, so the NCCL reductions of
            # block b overlap the backward of blocks b-1, b-2, ... (per-block checkpointing produces them block by block)
            def _reduce_when_ready(prm):
                pending.append(dist.all_reduce(prm.grad, async_op=True))
            if not getattr(args, "no_grad_sync", False):
                for p in params:
                    p.register_post_accumulate_grad_hook(_reduce_when_ready)

    if args.check:
        (y_a, g_a), (y_b, g_b) = iteration(True), iteration(False)

        def rel(a, b):
            return float((a.float() - b.float()).abs().max() / b.float().abs().max())

        def cos_sim(a, b):
            return float(F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0))
        chk = {"tool": "bench_denoiser --check", "model": args.model, **layers_desc, "img_tokens": n_img,
               "out_max_rel_diff_ours_vs_torch": round(rel(y_a, y_b), 5),
               "input_grad_cosine_ours_vs_torch": round(cos_sim(g_a, g_b), 6)}
        if emit:
            print(json.dumps(chk), flush=True)
        return chk

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    graph = None
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(3, args.warmup)):
                iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            iteration()
        graph.replay()
    else:
        for _ in range(args.warmup):
            iteration()
    sync_all()
    torch.cuda.reset_peak_memory_stats()
    L.profile_enable(not use_graph)  # per-kernel event timing records events at launch time: not inside a replayed graph
    e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0_.record()
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            iteration()
    e1_.record()
    sync_all()
    t_wall = time.perf_counter() - t_wall
    ms = torch.tensor([e0_.elapsed_time(e1_)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    attn_ms = sum(L.profile_read(k)[0] for k in (L.K_ATTN_FWD, L.K_ATTN_BWD, L.K_ATTN_BWD_DELTA, L.K_ATTN_BWD_DQ))
    attn_launches = sum(L.profile_read(k)[1] for k in (L.K_ATTN_FWD, L.K_ATTN_BWD))
    L.profile_enable(False)
    s_per_it = float(ms.item()) / 1e3 / args.steps
    line = None
    if rank == 0:
        line = {
            "tool": "bench_denoiser", "model": args.model, "arm": args.arm, "n_gpus": world,
            "it_per_s": round(1.0 / s_per_it, 5), "s_per_it": round(s_per_it, 4), "steps": args.steps, "warmup": args.warmup,
            "config": {**layers_desc, "img_tokens": n_img, "txt_tokens": n_txt, "hidden": C, "heads": heads,
                       "block_params": n_params, "trainable_params": sum(p.numel() for p in params),
                       "activation_checkpointing": ("per block, attention outputs kept (selective)" if keep_attn else ckpt), "dtype": "bf16", "optimizer": getattr(args, "optimizer", "none"),
                       "launch": "one CUDA graph per iteration" if use_graph else "eager",
                       "parallelism": "single" if world == 1 else f"ulysses_sp{world}",
                       "attention": ("b200vt tcgen05 kernels" if ours else (
                           "F.scaled_dot_product_attention (diffusers CogVideoXAttnProcessor2_0)" if args.model == "cogvideox"
                           else "flash_attn_varlen_func (FA2, the reference's mode=\"flash\")"))},
            "attention_algorithmic_tflop_per_it": round(attn_flops / 1e12, 1),
            "attention_executed_tflop_per_it": round(attn_exec / 1e12, 1),
            "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1),
            "wall_s": round(t_wall, 2),
        }
        if world > 1 and args.model == "wan":
            nosync = getattr(args, "no_grad_sync", False)
            line["grad_sync"] = ("skipped (--no-grad-sync diagnostic)" if nosync else
                                 f"all-reduce of {sum(p.numel() * p.element_size() for p in params) / 1e9:.2f} GB of replicated-parameter gradients per iteration")
        if ours and not use_graph:
            line["attention_kernel_s_per_it"] = round(attn_ms / 1e3 / args.steps, 4)
            line["attention_share_of_step"] = round(attn_ms / 1e3 / args.steps / s_per_it, 4)
            line["attention_launches_per_it"] = attn_launches / args.steps
            line["attention_tflops_in_step"] = round(attn_exec / world / (attn_ms / 1e3 / args.steps) / 1e12, 1)
        if emit:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if manage_dist:
            dist.destroy_process_group()
    return line


def main():
    run(parse())


if __name__ == "__main__":
    main()
