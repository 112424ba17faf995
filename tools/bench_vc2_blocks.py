"""Hot-path blocks of one VideoCrafter2 LoRA-finetune step (BASELINE.json configs[1]: bf16, batch 2, 320x512, 16 frames):
every SpatialTransformer (16), TemporalTransformer (17) and ResBlock (22) of the 3D-UNet, forward + backward, on synthetic
activations of each level's shape — the part of the UNet step this repository replaces (convolutions inside the ResBlocks
stay cuDNN in both arms; up/down-sampling, the temporal conv blocks and the time embedding are not part of it).

    python tools/bench_vc2_blocks.py [--arm ours|torch] [--steps K] [--warmup W] [--check]

arms
  ours    the reference constructors' shells (tests/helpers.py: same sub-module / parameter names) with the drop-in forwards
          of b200vt.blocks / functional (GroupNorm+SiLU, LayerNorm, attention kernels).
  torch   the same modules and weights driven by the reference's own op sequence, written out here: CrossAttention einsum /
          softmax / einsum (lvdm/modules/attention.py:101-170), BasicTransformerBlock (:299-310), SpatialTransformer
          (:376-392), TemporalTransformer (:475-519), ResBlock with GroupNormSpecific (fp32) + SiLU
          (networks/openaimodel3d.py:229-255, utils.py:192-203).
Both arms: bf16 autocast, frozen base weights with rank-4 adapters on to_q / to_k / to_v / to_out.0
(configs/001_videocrafter2/vc2_t2v_lora.yaml:9-11), no activation checkpointing. ResBlocks use the channel count of their
level (the skip-concatenated input widths of the decoder are not modelled). One iteration = every block once, forward +
backward, eager launches (so the host's launch path counts, as it does in the reference's training loop). CUDA events around
K iterations."""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys

import torch
import torch.nn.functional as F
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vt._lib as L  # noqa: E402

_spec = importlib.util.spec_from_file_location("b200vt_test_helpers", os.path.join(ROOT, "tests", "helpers.py"))
H = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(H)

BF16 = torch.bfloat16
BATCH, FRAMES, CTX_LEN, CTX_DIM, EMB = 2, 16, 77, 1024, 1280
# (channels, heads, h, w, spatial transformers, temporal transformers, resblocks)
LEVELS = [(320, 5, 40, 64, 5, 5, 5), (640, 10, 20, 32, 5, 5, 5), (1280, 20, 10, 16, 5, 5, 5), (1280, 20, 5, 8, 1, 1, 7)]


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


def add_lora(root: nn.Module) -> None:
    for p in root.parameters():
        p.requires_grad_(False)
    for m in root.modules():
        if isinstance(m, H.CrossAttentionShell):
            m.to_q, m.to_k, m.to_v = LoRALinear(m.to_q), LoRALinear(m.to_k), LoRALinear(m.to_v)
            m.to_out[0] = LoRALinear(m.to_out[0])


# ---- torch arm: the reference's op sequence ---------------------------------------------------------------------------
def attn_torch(m, x, context=None, mask=None):  # attention.py:101-170 (no relative position, no image tokens)
    h = m.heads
    q = m.to_q(x)
    context = x if context is None else context[:, : m.text_context_len, :]
    k, v = m.to_k(context), m.to_v(context)

    def split(t):
        b, n, _ = t.shape
        return t.view(b, n, h, -1).permute(0, 2, 1, 3).reshape(b * h, n, -1)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("b i d, b j d -> b i j", q, k) * m.scale
    if mask is not None:
        sim.masked_fill_(~(mask.repeat_interleave(h, dim=0) > 0.5), -torch.finfo(sim.dtype).max)
    sim = sim.softmax(dim=-1)
    out = torch.einsum("b i j, b j d -> b i d", sim, v)
    bh, n, d = out.shape
    out = out.view(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)
    return m.to_out(out)


def ff_torch(ff, x):  # FeedForward(glu=True) -> GEGLU (attention.py:522-548): chunk / gelu / mul as separate torch ops
    a, gate = ff.net[0].proj(x).chunk(2, dim=-1)
    return ff.net[2](ff.net[1](a * F.gelu(gate)))


def basic_torch(m, x, context=None, mask=None):  # attention.py:299-310
    x = attn_torch(m.attn1, m.norm1(x), context=None, mask=mask) + x
    x = attn_torch(m.attn2, m.norm2(x), context=context, mask=mask) + x
    return ff_torch(m.ff, m.norm3(x)) + x


def spatial_torch(m, x, context):  # attention.py:376-392 (use_linear)
    b, c, h, w = x.shape
    x_in = x
    x = m.norm(x)
    x = x.flatten(2).transpose(1, 2).contiguous()
    x = m.proj_in(x)
    for blk in m.transformer_blocks:
        x = basic_torch(blk, x, context=context)
    x = m.proj_out(x)
    x = x.transpose(1, 2).reshape(b, c, h, w).contiguous()
    return x + x_in


def temporal_torch(m, x):  # attention.py:475-519 (use_linear, only_self_att)
    b, c, t, h, w = x.shape
    x_in = x
    x = m.norm(x)
    x = x.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, t)  # (b h w) c t
    if m.use_linear:
        x = m.proj_in(x.transpose(1, 2).contiguous())
    else:  # Conv1d projections (the UNet's init_attn, openaimodel3d.py:418-432)
        x = m.proj_in(x).transpose(1, 2).contiguous()
    for blk in m.transformer_blocks:
        x = basic_torch(blk, x)
    if m.use_linear:
        x = m.proj_out(x).view(b, h, w, t, c).permute(0, 4, 3, 1, 2).contiguous()
    else:
        x = m.proj_out(x.transpose(1, 2).contiguous()).view(b, h, w, c, t).permute(0, 3, 4, 1, 2).contiguous()
    return x + x_in


def gn_specific(norm, x):  # utils.py:192-203: GroupNorm in fp32, result in x's dtype
    return F.group_norm(x.float(), norm.num_groups, norm.weight, norm.bias, norm.eps).type(x.dtype)


def resblock_torch(m, x, emb):  # openaimodel3d.py:229-255 (no up/down, no scale-shift norm, no temporal conv)
    h = m.in_layers[2](F.silu(gn_specific(m.in_layers[0], x)))
    emb_out = m.emb_layers(emb).type(h.dtype)[..., None, None]
    h = h + emb_out
    h = m.out_layers[3](m.out_layers[2](F.silu(gn_specific(m.out_layers[0], h))))
    return m.skip_connection(x) + h


# =====================================================================================================================
def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--arm", choices=("ours", "torch"), default="ours")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", action="store_true", help="both arms on the same weights: relative output differences")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    L.call("vt_init", 0)
    torch.manual_seed(20230211)
    g = torch.Generator(device=dev).manual_seed(20230211)

    def rn(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(BF16)

    work = []  # (kind, module, inputs (first one gets the gradient), output-gradient)
    torch.set_default_dtype(BF16)
    try:
        with torch.device(dev):
            for C, heads, h, w, n_st, n_tt, n_rb in LEVELS:
                xs = rn(BATCH * FRAMES, C, h, w)
                xt = rn(BATCH, C, FRAMES, h, w)
                ctx = rn(BATCH * FRAMES, CTX_LEN, CTX_DIM)
                emb = rn(BATCH * FRAMES, EMB)
                for _ in range(n_st):
                    work.append(("spatial", H.SpatialTransformerShell(C, heads, 64, depth=1, context_dim=CTX_DIM), (xs, ctx), rn(*xs.shape)))
                for _ in range(n_tt):
                    work.append(("temporal", H.TemporalTransformerShell(C, heads, 64, depth=1, temporal_length=FRAMES), (xt,), rn(*xt.shape)))
                for _ in range(n_rb):
                    work.append(("resblock", H.ResBlockShell(C, EMB, 0.0), (xs, emb), rn(*xs.shape)))
            # the UNet's first temporal transformer (init_attn): 8 heads x 64 on the 320-channel input, openaimodel3d.py:418-432
            C, _, h, w = LEVELS[0][:4]
            xt0 = rn(BATCH, C, FRAMES, h, w)
            work.append(("temporal", H.TemporalTransformerShell(C, 8, 64, depth=1, use_linear=False, temporal_length=FRAMES),
                         (xt0,), rn(*xt0.shape)))
    finally:
        torch.set_default_dtype(torch.float32)
    for _, m, _, _ in work:
        for n_, p_ in m.named_parameters():  # zero-initialised output projections re-drawn (trap 1)
            if n_.startswith("proj_out") or n_.startswith("out_layers.3"):
                nn.init.normal_(p_, std=0.02)
        add_lora(m)

    def run(kind, m, inputs, ours):
        x = inputs[0].detach().requires_grad_(True)
        with torch.autocast("cuda", dtype=BF16):
            if kind == "spatial":
                y = m(x, inputs[1]) if ours else spatial_torch(m, x, inputs[1])
            elif kind == "temporal":
                y = m(x) if ours else temporal_torch(m, x)
            else:
                y = m(x, inputs[1]) if ours else resblock_torch(m, x, inputs[1])
        return x, y

    if args.check:
        worst = {}
        for kind, m, inputs, dy in work[:: max(1, len(work) // 12)]:
            (xa, ya), (xb, yb) = run(kind, m, inputs, True), run(kind, m, inputs, False)
            ga, = torch.autograd.grad(ya, xa, dy)
            gb, = torch.autograd.grad(yb, xb, dy)
            e = float((ya.float() - yb.float()).abs().max() / yb.float().abs().max())
            c = float(F.cosine_similarity(ga.float().flatten(), gb.float().flatten(), dim=0))
            w_ = worst.setdefault(kind, [0.0, 1.0])
            w_[0], w_[1] = max(w_[0], e), min(w_[1], c)
        print(json.dumps({"tool": "bench_vc2_blocks --check", "max_rel_out_diff / min_input_grad_cosine (ours vs torch)":
                          {k: [round(v[0], 5), round(v[1], 6)] for k, v in worst.items()}}), flush=True)
        return

    ours = args.arm == "ours"
    params = [p for _, m, _, _ in work for p in m.parameters() if p.requires_grad]

    def iteration():
        for kind, m, inputs, dy in work:
            x, y = run(kind, m, inputs, ours)
            y.backward(dy)
        for p in params:
            p.grad = None

    for _ in range(args.warmup):
        iteration()
    torch.cuda.synchronize()
    per_kind = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    for kind in ("spatial", "temporal", "resblock"):  # per-kind split, one more pass
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k2, m, inputs, dy in work:
            if k2 == kind:
                x, y = run(k2, m, inputs, ours)
                y.backward(dy)
        b.record()
        torch.cuda.synchronize()
        per_kind[kind] = round(a.elapsed_time(b), 2)
    print(json.dumps({"tool": "bench_vc2_blocks", "arm": args.arm, "ms_per_pass": round(ms, 2), "steps": args.steps,
                      "blocks": {k: sum(1 for w_ in work if w_[0] == k) for k in ("spatial", "temporal", "resblock")},
                      "ms_by_kind": per_kind, "config": {"batch": BATCH, "frames": FRAMES, "levels": LEVELS, "dtype": "bf16",
                                                         "lora_rank": 4, "checkpointing": False}}), flush=True)


if __name__ == "__main__":
    main()
