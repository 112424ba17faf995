"""VideoCrafter2 T2V LoRA finetune STEP (BASELINE.json configs[1]; configs/001_videocrafter2/vc2_t2v_lora.yaml): the whole 3D-UNet
denoiser — every convolution, ResBlock (+ TemporalConvBlock), Spatial / TemporalTransformer, down / up-sampling, time and
fps embeddings, skip concatenations, output head — forward on noisy latents, MSE loss, backward, AdamW step on the rank-4
adapters, bf16 autocast, batch 2 x 16 frames x 320x512 (latent 2 x 4 x 16 x 40 x 64, 77 x 1024 text context).

    python tools/bench_vc2_unet.py [--arm ours|torch] [--steps K] [--warmup W] [--no-checkpoint] [--check]

The UNet is built here from the reference constructors' shells (tests/helpers.py) in the order UNetModel.__init__ builds
it (lvdm/modules/networks/openaimodel3d.py:313-640), with the reference's module names, so the reference's state dict loads
strictly (tests/test_tools_torch_arms.py pins the shell and the torch arm to the UNMODIFIED UNetModel on CPU). The GPU box
has no reference tree, hence the shell.
arms
  ours    drop-in forwards of b200vt.blocks / functional: GroupNorm+SiLU kernels (ResBlock, TemporalConvBlock, transformer
          norms), LayerNorm kernel, tcgen05 attention (spatial self / text cross), the one-warp temporal attention.
  torch   the same modules and weights through the reference's op sequence (tools/bench_vc2_blocks.py torch arm + the
          reference's plain nn.Sequential TemporalConvBlock), i.e. what the unpatched reference launches.
Both arms: per-block activation checkpointing as the config sets (`use_checkpoint: true`; lvdm/modules/utils.py:112-125),
frozen base weights, LoRA on to_q / to_k / to_v (target_modules of the config), dropout 0.1 inside the temporal conv blocks
(training mode), eager launches, CUDA events around K steps."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import torch
import torch.nn.functional as F
from torch import nn
from torch.utils.checkpoint import checkpoint

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_vc2_blocks as VB  # noqa: E402  (torch-arm op sequences, LoRALinear, shells module H)

H = VB.H
BF16 = torch.bfloat16
VC2 = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1), num_res_blocks=2,
           channel_mult=(1, 2, 4, 4), num_head_channels=64, transformer_depth=1, context_dim=1024, temporal_length=16,
           temporal_conv=True, addition_attention=True, fps_cond=True)


class Downsample(nn.Module):  # openaimodel3d.py:72-101 (use_conv, dims=2)
    def __init__(self, ch):
        super().__init__()
        self.op = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class Upsample(nn.Module):  # openaimodel3d.py:104-136
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


def timestep_embedding(t, dim, max_period=10000):  # lvdm/modules/utils.py timestep_embedding (sinusoidal, cos first)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class VC2UNet(nn.Module):
    """UNetModel (openaimodel3d.py:313-706) for configurations of the VideoCrafter2 family: dims = 2, conv resampling,
    use_linear transformers, temporal self-attention without relative position, optional temporal conv / init_attn / fps."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, channel_mult,
                 num_head_channels, transformer_depth, context_dim, temporal_length, temporal_conv=True,
                 addition_attention=True, fps_cond=True):
        super().__init__()
        self.model_channels, self.fps_cond, self.addition_attention = model_channels, fps_cond, addition_attention
        emb_dim = model_channels * 4

        def embed():
            return nn.Sequential(nn.Linear(model_channels, emb_dim), nn.SiLU(), nn.Linear(emb_dim, emb_dim))

        def res(cin, cout):
            return H.ResBlockShell(cin, emb_dim, 0.0, out_channels=cout, use_temporal_conv=temporal_conv)

        def attn_layers(ch):
            heads = ch // num_head_channels
            return [H.SpatialTransformerShell(ch, heads, num_head_channels, depth=transformer_depth, context_dim=context_dim),
                    H.TemporalTransformerShell(ch, heads, num_head_channels, depth=transformer_depth,
                                               temporal_length=temporal_length)]

        self.time_embed = embed()
        if fps_cond:
            self.fps_embedding = embed()
        self.input_blocks = nn.ModuleList([nn.Sequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        if addition_attention:
            self.init_attn = nn.Sequential(H.TemporalTransformerShell(model_channels, 8, num_head_channels, use_linear=False,
                                                                      depth=transformer_depth, temporal_length=temporal_length))
        chans, ch, ds = [model_channels], model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers += attn_layers(ch)
                self.input_blocks.append(nn.Sequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(nn.Sequential(Downsample(ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = nn.Sequential(res(ch, ch), *attn_layers(ch), res(ch, ch))
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers += attn_layers(ch)
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch))
                    ds //= 2
                self.output_blocks.append(nn.Sequential(*layers))
        self.out = nn.Sequential(nn.GroupNorm(32, ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))

    # ---- one TimestepEmbedSequential (openaimodel3d.py:35-57) ----------------------------------------------------
    def _run(self, seq, h, emb, context, b, ours, ckpt):
        for layer in seq:
            if isinstance(layer, H.ResBlockShell):
                fn = (lambda x, e, _m=layer: _m(x, e, b)) if ours else (lambda x, e, _m=layer: resblock_torch(_m, x, e, b))
                h = checkpoint(fn, h, emb, use_reentrant=False) if ckpt else fn(h, emb)
            elif isinstance(layer, H.SpatialTransformerShell):
                h = spatial(layer, h, context, ours, ckpt)
            elif isinstance(layer, H.TemporalTransformerShell):
                bt, c, hh, ww = h.shape
                x5 = h.view(b, bt // b, c, hh, ww).permute(0, 2, 1, 3, 4)  # (b f) c h w -> b c f h w
                x5 = temporal(layer, x5, ours, ckpt)
                h = x5.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)
            else:
                h = layer(h)
        return h

    def forward(self, x, timesteps, context, fps=24, ours=True, ckpt=True):
        for m in self.modules():
            if isinstance(m, H.BasicBlockShell):
                m.checkpoint = ckpt
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels).to(x.dtype))
        if self.fps_cond:
            fps_t = torch.full_like(timesteps, fps) if isinstance(fps, int) else fps
            emb = emb + self.fps_embedding(timestep_embedding(fps_t, self.model_channels).to(x.dtype))
        b, _, t, hh, ww = x.shape
        context = context.repeat_interleave(repeats=t, dim=0)
        emb = emb.repeat_interleave(repeats=t, dim=0)
        h = x.permute(0, 2, 1, 3, 4).reshape(b * t, -1, hh, ww)
        hs = []
        for i, seq in enumerate(self.input_blocks):
            h = self._run(seq, h, emb, context, b, ours, ckpt)
            if i == 0 and self.addition_attention:
                h = self._run(self.init_attn, h, emb, context, b, ours, ckpt)
            hs.append(h)
        h = self._run(self.middle_block, h, emb, context, b, ours, ckpt)
        for seq in self.output_blocks:
            h = self._run(seq, torch.cat([h, hs.pop()], dim=1), emb, context, b, ours, ckpt)
        if ours:
            import b200vt.blocks as Bk
            y = self.out[2](Bk._lvdm_gn(self.out[0], h, silu=True))
        else:
            y = self.out[2](F.silu(VB.gn_specific(self.out[0], h)))
        return y.view(b, t, -1, hh, ww).permute(0, 2, 1, 3, 4)


# ---- torch arm of the transformer wrappers: the checkpoint sits on BasicTransformerBlock.forward (attention.py:283-297);
# the `ours` arm calls the shells, i.e. the drop-in forwards patch_blocks() installs, whose BasicBlockShell checkpoints alike
def _blocks_torch(m, x, context, ckpt):
    for blk in m.transformer_blocks:
        fn = (lambda t, c, _b=blk: VB.basic_torch(_b, t, context=c)) if context is not None else (lambda t, _b=blk: VB.basic_torch(_b, t))
        args = (x, context) if context is not None else (x,)
        x = checkpoint(fn, *args, use_reentrant=False) if ckpt else fn(*args)
    return x


def spatial(m, x, context, ours, ckpt):  # attention.py:376-392 (use_linear)
    if ours:
        return m(x, context)
    b, c, h, w = x.shape
    t = m.proj_in(m.norm(x).flatten(2).transpose(1, 2).contiguous())
    t = m.proj_out(_blocks_torch(m, t, context, ckpt))
    return t.transpose(1, 2).reshape(b, c, h, w).contiguous() + x


def temporal(m, x, ours, ckpt):  # attention.py:475-519 (use_linear, only_self_att)
    if ours:
        return m(x)
    b, c, t, h, w = x.shape
    y = m.norm(x)
    if m.use_linear:
        y = m.proj_in(y.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c))
        y = m.proj_out(_blocks_torch(m, y, None, ckpt))
        return y.view(b, h, w, t, c).permute(0, 4, 3, 1, 2).contiguous() + x
    y = m.proj_in(y.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, t)).transpose(1, 2).contiguous()  # Conv1d over (bhw, c, t)
    y = m.proj_out(_blocks_torch(m, y, None, ckpt).transpose(1, 2).contiguous())
    return y.view(b, h, w, c, t).permute(0, 3, 4, 1, 2).contiguous() + x


def resblock_torch(m, x, emb, batch_size):  # openaimodel3d.py:229-255 incl. the temporal conv block (:248-253)
    h = VB.resblock_torch(m, x, emb)
    if m.use_temporal_conv and batch_size:
        bt, ch, hh, ww = h.shape
        h5 = h.view(batch_size, bt // batch_size, ch, hh, ww).transpose(1, 2)
        tc = m.temopral_conv
        y = h5
        for stage in (tc.conv1, tc.conv2, tc.conv3, tc.conv4):  # TemporalConvBlock.forward (:303-310): plain Sequentials
            y = stage(y)
        h = (y + h5).transpose(1, 2).reshape(bt, ch, hh, ww)
    return h


def add_lora_qkv(root: nn.Module) -> None:
    """peft target_modules ["to_q", "to_k", "to_v"] (vc2_t2v_lora.yaml:9), rank 4, alpha 1; base weights frozen."""
    for p in root.parameters():
        p.requires_grad_(False)
    for m in root.modules():
        if isinstance(m, H.CrossAttentionShell):
            m.to_q, m.to_k, m.to_v = VB.LoRALinear(m.to_q), VB.LoRALinear(m.to_k), VB.LoRALinear(m.to_v)


def build(dev, cfg=None, dtype=torch.float32):
    torch.manual_seed(20230211)
    with torch.device(dev):
        net = VC2UNet(**(cfg or VC2))
    for n_, p_ in net.named_parameters():  # zero-initialised layers re-drawn so that both passes are live (SURVEY §4 trap 1)
        if float(p_.detach().abs().max()) == 0.0 or n_.endswith("proj_out.weight") or n_.endswith("conv4.3.weight"):
            nn.init.normal_(p_, std=0.02)
    return net.to(dtype)


def run(args, emit=True):
    import b200vt._lib as L
    dev = torch.device("cuda", torch.cuda.current_device())
    L.call("vt_init", dev.index)
    net = build(dev)
    add_lora_qkv(net)
    net.train()
    if args.arm == "ours" and not getattr(args, "nchw", False):
        import b200vt.patch as P
        P.lvdm_channels_last(net)  # channels-last activation flow of the drop-ins (the reference arm stays NCHW, as it is)
    frozen_cast = 0
    if args.arm == "ours" and getattr(args, "frozen_bf16", False):
        import b200vt.patch as P
        frozen_cast = P.cast_frozen_weights(net)  # bit-identical under autocast; no per-step weight casts
    params = [p for p in net.parameters() if p.requires_grad]
    g = torch.Generator(device=dev).manual_seed(20230211)
    B, T = 2, 16
    x0 = torch.randn(B, 4, T, 40, 64, device=dev, generator=g)
    noise = torch.randn(B, 4, T, 40, 64, device=dev, generator=g)
    ctx = torch.randn(B, 77, 1024, device=dev, generator=g)
    tt = torch.randint(0, 1000, (B,), device=dev, generator=g)
    ckpt = not args.no_checkpoint
    use_graph = bool(getattr(args, "graph", False))
    if use_graph:
        ckpt = False  # whole-step CUDA graph: no activation checkpointing (its RNG-state save is a host operation)
    opt = torch.optim.AdamW(params, lr=6e-6, fused=True, capturable=use_graph)

    def step(ours, zero=True):
        x = (0.7 * x0 + 0.7 * noise).requires_grad_(False)
        with torch.autocast("cuda", dtype=BF16):
            pred = net(x, tt, ctx, 24, ours=ours, ckpt=ckpt)
            loss = F.mse_loss(pred.float(), noise)
        loss.backward()
        opt.step()
        if zero:
            opt.zero_grad(set_to_none=True)
        return loss

    if args.check:
        net.eval()  # dropout off: both arms must agree
        outs = {}
        for ours in (True, False):
            x = (0.7 * x0 + 0.7 * noise)
            with torch.autocast("cuda", dtype=BF16):
                pred = net(x, tt, ctx, 24, ours=ours, ckpt=False)
                loss = F.mse_loss(pred.float(), noise)
            loss.backward()
            gsel = torch.cat([p.grad.float().flatten() for p in params[:24]])
            outs[ours] = (pred.detach().float(), gsel.clone())
            for p in params:
                p.grad = None
        rel = float((outs[True][0] - outs[False][0]).abs().max() / outs[False][0].abs().max())
        cosg = float(F.cosine_similarity(outs[True][1], outs[False][1], dim=0))
        line = {"tool": "bench_vc2_unet --check", "out_max_rel_diff_ours_vs_torch": round(rel, 5),
                "lora_grad_cosine_ours_vs_torch": round(cosg, 6)}
        if emit:
            print(json.dumps(line), flush=True)
        return line

    ours = args.arm == "ours"
    graph = None
    if use_graph:
        # The whole training step — forward, loss, backward, fused AdamW — captured ONCE into a CUDA graph and replayed:
        # the ~10 000 eager launches of a step (the B200 finishes most of them faster than the host can issue the next) become
        # one graph launch. Static inputs, graph-safe philox for the dropout masks, capturable optimizer.
        import b200vt.graph as G
        def graphed():
            opt.zero_grad(set_to_none=True)  # gradients are None when the capture starts: the captured backward ASSIGNS them
            return step(ours, zero=False)

        graph = G.GraphedStep(graphed, warmup=max(3, args.warmup))
        static_loss = graph.result
        for _ in range(2):
            graph()
    else:
        for _ in range(args.warmup):
            step(ours)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        if graph is not None:
            graph()
            loss = static_loss
        else:
            loss = step(ours)
    e1.record()
    torch.cuda.synchronize()
    s_per_it = e0.elapsed_time(e1) / 1e3 / args.steps
    line = {"tool": "bench_vc2_unet", "arm": args.arm, "s_per_it": round(s_per_it, 4), "it_per_s": round(1.0 / s_per_it, 3),
            "steps": args.steps, "warmup": args.warmup, "loss": round(float(loss.detach()), 4),
            "config": {"batch": B, "frames": T, "latent": [4, T, 40, 64], "context": [77, 1024], "dtype": "bf16 autocast",
                       "unet_params": sum(p.numel() for p in net.parameters()), "trainable_params": sum(p.numel() for p in params),
                       "lora": "rank 4 on to_q/to_k/to_v", "activation_checkpointing": ckpt, "optimizer": "AdamW (fused)",
                       "launch": "one CUDA graph per step" if use_graph else "eager",
                       "activation_layout": ("channels_last" if args.arm == "ours" and not getattr(args, "nchw", False) else "nchw"),
                       "frozen_weights": "bf16 (cast once; %d tensors)" % frozen_cast if frozen_cast else "fp32, cast by autocast at every use"},
            "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
    if emit:
        print(json.dumps(line), flush=True)
    return line


def parse(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--arm", choices=("ours", "torch"), default="ours")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-checkpoint", action="store_true")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--nchw", action="store_true", help="ours arm without the channels-last activation flow")
    ap.add_argument("--frozen-bf16", action="store_true",
                    help="ours arm: store the frozen Conv / Linear weights in bf16 once (patch.cast_frozen_weights; bit-identical "
                    "under autocast)")
    ap.add_argument("--graph", action="store_true",
                    help="capture the whole step (forward, loss, backward, AdamW) into one CUDA graph and replay it (implies "
                    "--no-checkpoint)")
    return ap.parse_args(argv)


if __name__ == "__main__":
    run(parse())
