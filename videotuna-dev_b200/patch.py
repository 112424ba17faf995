"""patch_videotuna(): install the b200vt hot path underneath an importable VideoTuna tree.

Nothing here re-implements VideoTuna: each hook is a place where the reference itself already swaps implementations
(instance-level `self.forward = self.efficient_forward`, name imports of `attention` / `flash_attention`, the
`hybrid_seq_parallel_attn` attribute). A patched callable first tries the CUDA path; `functional.Unsupported` (fp32 or
CPU tensors, masks, dropout, causal, unsupported head dims ...) routes the call to the ORIGINAL reference callable,
never to a CPU kernel of ours.

    import b200vt
    b200vt.patch.patch_videotuna()          # before building the model; idempotent
    b200vt.patch.patch_sp()                 # sequence parallelism without xfuser: the reference flows enter SP unmodified
    b200vt.patch.install_ulysses(dit)       # or by hand: every block's hybrid_seq_parallel_attn = UlyssesAttention()
"""
from __future__ import annotations

import functools
import importlib
import types
from typing import Callable, Dict, Optional

import torch

from . import blocks as Bk
from . import functional as Fn
from . import sp
from . import xfuser_shim

_ORIGINALS: Dict[str, Callable] = {}


def _wrap(name: str, original: Callable, fast: Callable) -> Callable:
    """fast(*a, **k) with fallback to the untouched reference callable on Unsupported."""
    if getattr(original, "_b200vt_patched", False):
        return original
    _ORIGINALS[name] = original

    @functools.wraps(original)
    def patched(*args, **kwargs):
        try:
            return fast(*args, **kwargs)
        except Fn.Unsupported:
            return original(*args, **kwargs)

    patched._b200vt_patched = True
    patched._b200vt_original = original
    return patched


def _try_import(name: str):
    try:
        return importlib.import_module(name)
    except Exception:  # noqa: BLE001  (a missing optional dependency of the reference must not break the others)
        return None


def patch_lvdm(module=None, xformers_too: bool = True) -> bool:
    """CrossAttention.forward (videotuna/models/lvdm/modules/attention.py:101-170) -> functional.lvdm_cross_attention_forward,
    installed on the class. With xformers present the reference's spatial instances switch themselves to
    `efficient_forward` at construction (`self.forward = self.efficient_forward`, :98-99): xformers_too=True also replaces
    the class's `efficient_forward` (same signature; `Unsupported` inputs reach the original xformers body), so instances
    built AFTER the patch bind ours; for models built before it call rebind_lvdm_instances(model)."""
    mod = module or _try_import("videotuna.models.lvdm.modules.attention")
    if mod is None:
        return False
    cls = mod.CrossAttention
    cls.forward = _wrap("lvdm.CrossAttention.forward", cls.forward, Fn.lvdm_cross_attention_forward)
    if xformers_too and hasattr(cls, "efficient_forward"):
        cls.efficient_forward = _wrap("lvdm.CrossAttention.efficient_forward", cls.efficient_forward,
                                      Fn.lvdm_cross_attention_forward)
    return True


def rebind_lvdm_instances(model: torch.nn.Module) -> int:
    """Drop the instance-level `forward` that reference CrossAttention modules built before patch_lvdm() set on themselves
    (the xformers switch, attention.py:98-99), so that the patched class attribute applies. Returns how many were reset."""
    n = 0
    for m in model.modules():
        if type(m).__name__ == "CrossAttention" and "forward" in vars(m):
            del vars(m)["forward"]
            n += 1
    return n


def patch_hunyuan(modules=None) -> int:
    """Rebind the name `attention` that models.py / token_refiner.py imported from attenion.py
    (hyvideo_t2v/modules/models.py:14, token_refiner.py:8; the i2v package is a byte-identical twin) to
    functional.hunyuan_attention (same signature, attenion.py:60-156)."""
    names = modules or [f"videotuna.models.hunyuan.{pkg}.modules.{m}" for pkg in ("hyvideo_t2v", "hyvideo_i2v")
                        for m in ("models", "token_refiner", "attenion")]
    n = 0
    for name in names:
        mod = name if isinstance(name, types.ModuleType) else _try_import(name)
        if mod is None or not hasattr(mod, "attention"):
            continue
        mod.attention = _wrap(f"{mod.__name__}.attention", mod.attention, Fn.hunyuan_attention)
        n += 1
    return n


def patch_wan(modules=None) -> int:
    """Rebind `flash_attention` in wan/modules/model.py (imported at :10, called by WanSelfAttention :146 and the
    cross-attention classes) and in wan/modules/attention.py (used by attention() :133-179)."""
    names = modules or ["videotuna.models.wan.wan.modules.model", "videotuna.models.wan.wan.modules.attention"]
    n = 0
    for name in names:
        mod = name if isinstance(name, types.ModuleType) else _try_import(name)
        if mod is None or not hasattr(mod, "flash_attention"):
            continue
        mod.flash_attention = _wrap(f"{mod.__name__}.flash_attention", mod.flash_attention, Fn.wan_flash_attention)
        n += 1
    return n


def _patch_method(mod, cls_name: str, attr: str, fast: Callable) -> int:
    cls = getattr(mod, cls_name, None)
    if cls is None or not hasattr(cls, attr):
        return 0
    setattr(cls, attr, _wrap(f"{mod.__name__}.{cls_name}.{attr}", getattr(cls, attr), fast))
    return 1


def patch_blocks(lvdm: bool = True, hunyuan: bool = True, wan: bool = True) -> Dict[str, int]:
    """Replace the block-level forwards with the fused versions in b200vt.blocks.
    Independent of patch_videotuna(): the fused blocks call b200vt's attention functions directly. Idempotent."""
    done = {"lvdm": 0, "hunyuan": 0, "wan": 0}
    if lvdm:
        mod = _try_import("videotuna.models.lvdm.modules.attention")
        if mod is not None:
            done["lvdm"] += _patch_method(mod, "BasicTransformerBlock", "_forward", Bk.lvdm_basic_block_forward)
            done["lvdm"] += _patch_method(mod, "SpatialTransformer", "forward", Bk.lvdm_spatial_transformer_forward)
            done["lvdm"] += _patch_method(mod, "TemporalTransformer", "forward", Bk.lvdm_temporal_transformer_forward)
            done["lvdm"] += _patch_method(mod, "GEGLU", "forward", Fn.lvdm_geglu_forward)
        net = _try_import("videotuna.models.lvdm.modules.networks.openaimodel3d")
        if net is not None:
            done["lvdm"] += _patch_method(net, "ResBlock", "_forward", Bk.lvdm_resblock_forward)
            done["lvdm"] += _patch_method(net, "TemporalConvBlock", "forward", Bk.lvdm_temporal_conv_block_forward)
    if hunyuan:
        # The drop-ins take the i2v twins' full signature (condition_type, token_replace_vec, frist_frame_token_num:
        # hyvideo_i2v/modules/models.py:136-149, 371-384) — HunyuanVideoFlow drives T2V and I2V through the i2v DiT, which
        # always passes all 11 arguments positionally (:749-761, 776-788).
        for pkg in ("hyvideo_t2v", "hyvideo_i2v"):
            mod = _try_import(f"videotuna.models.hunyuan.{pkg}.modules.models")
            if mod is None:
                continue
            done["hunyuan"] += _patch_method(mod, "MMDoubleStreamBlock", "forward", Bk.hunyuan_double_block_forward)
            done["hunyuan"] += _patch_method(mod, "MMSingleStreamBlock", "forward", Bk.hunyuan_single_block_forward)
            if hasattr(mod, "parallel_attention"):
                mod.parallel_attention = _wrap(f"{mod.__name__}.parallel_attention", mod.parallel_attention,
                                               Bk.hunyuan_parallel_attention)
                done["hunyuan"] += 1
    if wan:
        mod = _try_import("videotuna.models.wan.wan.modules.model")
        if mod is not None:
            done["wan"] += _patch_method(mod, "WanSelfAttention", "forward", Bk.wan_self_attention_forward)
            done["wan"] += _patch_method(mod, "WanT2VCrossAttention", "forward", Bk.wan_t2v_cross_attention_forward)
            done["wan"] += _patch_method(mod, "WanI2VCrossAttention", "forward", Bk.wan_i2v_cross_attention_forward)
            done["wan"] += _patch_method(mod, "WanAttentionBlock", "forward", Bk.wan_attention_block_forward)
    return done


def lvdm_channels_last(model: torch.nn.Module) -> int:
    """Opt an lvdm 3D-UNet (openaimodel3d.UNetModel) into the channels-last activation flow of the block drop-ins: the
    weights of every Conv2d / Conv3d are re-laid ONCE as torch.channels_last / channels_last_3d (values, names and
    state-dict loading are unchanged). The ResBlock drop-in then moves the activation to channels-last at its first call,
    GroupNorm + SiLU runs in the channels-last kernels, cuDNN's tensor-core convolutions need no nchw <-> nhwc conversion
    kernels around them (10.8 % of the VideoCrafter2 LoRA step), and SpatialTransformer's `b c h w <-> b (h w) c` become
    views. Call after patch_blocks() and after the model is on the GPU; returns the number of convolutions converted."""
    n = 0
    for m in model.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.to(memory_format=torch.channels_last)
            n += 1
        elif isinstance(m, torch.nn.Conv3d):
            m.to(memory_format=torch.channels_last_3d)
            n += 1
    return n


def cast_frozen_weights(model: torch.nn.Module, dtype: torch.dtype = torch.bfloat16) -> int:
    """LoRA finetuning under autocast (vc2_t2v_lora.yaml:125 `precision: bf16`; lvdm/ddpm3d.py:112-117 freezes everything but
    the adapters): store the FROZEN Conv / Linear weights and biases in the autocast dtype once. Autocast rounds them to that
    dtype at every use anyway (its cast cache only holds tensors that require grad), so the results are bit-identical; what
    goes away is one fp32 -> bf16 cast kernel per layer per step (1 342 launches, 3.4 % of the VideoCrafter2 LoRA step) and
    half of the weights' memory. Normalisation layers (which autocast runs in fp32) and every trainable tensor are left
    alone. Opt-in: the state dict then holds `dtype` weights. Returns the number of tensors converted."""
    n = 0
    for m in model.modules():
        if isinstance(m, (torch.nn.Linear, torch.nn.Conv1d, torch.nn.Conv2d, torch.nn.Conv3d)):
            for name in ("weight", "bias"):
                p = getattr(m, name, None)
                if isinstance(p, torch.nn.Parameter) and not p.requires_grad and p.dtype == torch.float32:
                    p.data = p.data.to(dtype)  # keeps the memory format (channels-last weights stay channels-last)
                    n += 1
    return n


def set_diffusers_processors(transformer: torch.nn.Module) -> int:
    """diffusers models (CogVideoXTransformer3DModel, HunyuanVideoTransformer3DModel; reference call sites
    cogvideo_hf/cogvideo_pl.py:123 and hyvideo_t2v/hunyuanvideo.py:209): install the duck-typed processors on every
    `Attention` module whose current processor is the stock CogVideoX / HunyuanVideo one. Returns how many were set."""
    n = 0
    for m in transformer.modules():
        proc = getattr(m, "processor", None)
        if proc is None or not hasattr(m, "set_processor"):
            continue
        name = type(proc).__name__
        if name.startswith("CogVideoXAttnProcessor"):
            m.set_processor(_FallbackProcessor(Bk.CogVideoXAttnProcessor(), proc))
            n += 1
        elif name.startswith("HunyuanVideoAttnProcessor"):
            m.set_processor(_FallbackProcessor(Bk.HunyuanVideoAttnProcessor(), proc))
            n += 1
    return n


def set_diffusers_blocks(transformer: torch.nn.Module) -> int:
    """Bind `blocks.cogvideox_block_forward` on every `CogVideoXBlock` of a diffusers CogVideoXTransformer3DModel
    (instance-level, like the reference's own `self.forward = self.efficient_forward` swap, lvdm attention.py:98-99);
    `Unsupported` inputs go to the block's stock forward. Call set_diffusers_processors() as well for the attention."""
    import types
    n = 0
    for m in transformer.modules():
        if type(m).__name__ != "CogVideoXBlock" or getattr(m, "_b200vt_stock_forward", None) is not None:
            continue
        stock = m.forward

        def forward(self, *args, _stock=stock, **kwargs):
            try:
                return Bk.cogvideox_block_forward(self, *args, **kwargs)
            except Fn.Unsupported:
                return _stock(*args, **kwargs)

        m._b200vt_stock_forward = stock
        m.forward = types.MethodType(forward, m)
        n += 1
    return n


class _FallbackProcessor:
    """Try the CUDA processor; `Unsupported` (masks, fp32, CPU, ...) goes to the stock processor it replaced."""

    def __init__(self, fast, original):
        self.fast, self.original = fast, original

    def __call__(self, attn, *args, **kwargs):
        try:
            return self.fast(attn, *args, **kwargs)
        except Fn.Unsupported:
            return self.original(attn, *args, **kwargs)


def patch_videotuna(lvdm: bool = True, hunyuan: bool = True, wan: bool = True, blocks: bool = False) -> Dict[str, int]:
    """Install every hook whose reference module imports in this environment; returns what was patched.
    blocks=True additionally installs the fused block forwards (patch_blocks)."""
    done = {"lvdm": 0, "hunyuan": 0, "wan": 0}
    if lvdm:
        done["lvdm"] = int(patch_lvdm())
    if hunyuan:
        done["hunyuan"] = patch_hunyuan()
    if wan:
        done["wan"] = patch_wan()
    if blocks:
        for k, v in patch_blocks(lvdm, hunyuan, wan).items():
            done[k] += v
    return done


def unpatch_videotuna() -> None:
    """Restore every original callable (tests)."""
    for name, original in list(_ORIGINALS.items()):
        if name in ("lvdm.CrossAttention.forward", "lvdm.CrossAttention.efficient_forward"):
            mod = _try_import("videotuna.models.lvdm.modules.attention")
            if mod is not None:
                setattr(mod.CrossAttention, name.rsplit(".", 1)[1], original)
        else:
            modname, attr = name.rsplit(".", 1)
            mod = _try_import(modname)
            if mod is not None:
                setattr(mod, attr, original)
            else:  # "<module>.<Class>.<method>" (patch_blocks)
                modname, cls_name = modname.rsplit(".", 1)
                mod = _try_import(modname)
                if mod is not None and hasattr(mod, cls_name):
                    setattr(getattr(mod, cls_name), attr, original)
        _ORIGINALS.pop(name, None)


# ---------------------------------------------------------------------------------------------------------------------
# sequence parallelism
# ---------------------------------------------------------------------------------------------------------------------
def install_ulysses(transformer: torch.nn.Module, group=None) -> int:
    """What parallelize_transformer does with xfuser (videotuna/flow/hunyuanvideo.py:154-157): every double/single
    stream block gets `hybrid_seq_parallel_attn`, here a b200vt.sp.UlyssesAttention with xfuser's call signature."""
    attn = sp.UlyssesAttention(group)
    n = 0
    for name in ("double_blocks", "single_blocks"):
        for block in getattr(transformer, name, []):
            block.hybrid_seq_parallel_attn = attn
            n += 1
    return n


def wan_sp_rope_tables(grid_size, freqs: torch.Tensor, s_local: int, group=None):
    """cos/sin tables (s_local, D) for this rank's slice of the token sequence, equal to what the reference's
    sequence-parallel rope_apply multiplies by (xdit_context_parallel.py:26-63): the (f*h*w, D/2) complex table is
    padded with ones to s_local * P rows (pad_freqs :12-22) and rows [rank*s_local, (rank+1)*s_local) are taken."""
    f, h, w = (int(v) for v in grid_size)
    c = freqs.shape[1]
    fs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    fr = torch.cat([fs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), fs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                    fs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, -1)
    P, r = sp._world(group), sp._rank(group)
    pad = s_local * P - fr.shape[0]
    if pad > 0:
        fr = torch.cat([fr, torch.ones(pad, c, dtype=fr.dtype, device=fr.device)], dim=0)
    fr = fr[r * s_local:(r + 1) * s_local]
    return (fr.real.float().repeat_interleave(2, dim=1).contiguous(),
            fr.imag.float().repeat_interleave(2, dim=1).contiguous())


_WAN_SP_ROPE_CACHE: dict = {}


def _wan_sp_rope_cached(grid, freqs: torch.Tensor, s_local: int, group, device):
    key = (tuple(int(v) for v in grid), freqs.data_ptr(), s_local, id(group), sp._rank(group), str(device))
    hit = _WAN_SP_ROPE_CACHE.get(key)
    if hit is None:
        cos, sin = wan_sp_rope_tables(grid, freqs, s_local, group)
        hit = (cos.to(device), sin.to(device))
        if len(_WAN_SP_ROPE_CACHE) > 16:
            _WAN_SP_ROPE_CACHE.clear()
        _WAN_SP_ROPE_CACHE[key] = hit
    return hit


def wan_usp_attn_forward(self, x, seq_lens, grid_sizes, freqs, dtype=torch.bfloat16, group=None):
    """Replacement for usp_attn_forward (xdit_context_parallel.py:149-192), bound with types.MethodType onto
    WanSelfAttention like the reference does (wan/text2video.py:261-271): the q/k/v/o projections stay the module's own
    layers; the full-dim RMSNorm and the RoPE of this rank's token slice run as one bf16 pass per tensor (the same fused
    kernel as the single-GPU block, with the per-rank table slice); attention is Ulysses."""
    b, s, n, d = *x.shape[:2], self.num_heads, self.head_dim
    if not (x.is_cuda and b == 1):
        raise Fn.Unsupported("sequence-parallel Wan attention needs CUDA tensors and batch 1 (as the reference pipeline)")
    grid = grid_sizes[0].tolist() if isinstance(grid_sizes, torch.Tensor) else list(grid_sizes[0])
    cos, sin = _wan_sp_rope_cached(grid, freqs, s, group, x.device)

    def proj(lin):
        t = lin(x)
        return (t if t.dtype == torch.bfloat16 else t.to(dtype)).view(b, s, n, d)

    def norm_rope(t, norm):
        w, eps = Bk._rms_weight(norm)
        return Fn.qk_rmsnorm_rope(t, w, cos, sin, per_head=False, eps=eps if w is not None else 1e-6)

    q = norm_rope(proj(self.q), self.norm_q)
    k = norm_rope(proj(self.k), self.norm_k)
    out = sp.UlyssesAttention(group)(None, q, k, proj(self.v), window_size=self.window_size)
    return self.o(out.flatten(2).to(x.dtype))


def wan_usp_dit_forward(self, x, t, context, seq_len, clip_fea=None, y=None, group=None):
    """xfuser-free replacement for usp_dit_forward (xdit_context_parallel.py:66-146), bound onto a WanModel with
    types.MethodType like the reference does (wan/text2video.py:270). The reference's wrapper restates WanModel.forward
    (wan/modules/model.py:482-574) with two changes — `torch.chunk(x, P, dim=1)[rank]` on the padded token sequence before
    the first block (:121-127) and `get_sp_group().all_gather(x, dim=1)` after the head (:142). Here the model's OWN class
    forward runs and those two steps are attached where they belong: a pre-hook on the first block keeps this rank's token
    chunk, a hook on the head gathers the chunks. Patch embedding, time / text / CLIP embeddings and unpatchify are
    therefore exactly the model's, whatever its version."""
    if sp._world(group) == 1:
        return type(self).forward(self, x, t, context, seq_len, clip_fea=clip_fea, y=y)

    def keep_my_chunk(_module, args, kwargs):
        return (sp.shard_sequence(args[0], dim=1, group=group), *args[1:]), kwargs

    def gather_chunks(_module, _args, out):
        return sp.gather_sequence(out, dim=1, group=group)

    hooks = [self.blocks[0].register_forward_pre_hook(keep_my_chunk, with_kwargs=True),
             self.head.register_forward_hook(gather_chunks)]
    try:
        return type(self).forward(self, x, t, context, seq_len, clip_fea=clip_fea, y=y)
    finally:
        for h in hooks:
            h.remove()


def hunyuan_parallelize_transformer(pipe, group=None, attn_fn: Optional[Callable] = None):
    """xfuser-free replacement for parallelize_transformer (flow/hunyuanvideo.py:114-178; twin in
    hyvideo_t2v/inference.py:48-110): wraps `pipe.transformer.forward` so that every rank denoises one slab of the latent.
    The latent (B, C, T, H, W) is cut along the patch rows when H/2 divides by the SP size, else along the patch columns;
    the RoPE tables, laid out as (T, H/2, W/2) tokens, are cut the same way; every double / single stream block gets the
    Ulysses attention object; the output slabs are gathered back along the cut axis. attn_fn: attention core for
    sp.UlyssesAttention (default: the CUDA kernels; the CPU tests of the exchange logic inject the oracle)."""
    transformer = pipe.transformer
    inner = transformer.forward
    attn = sp.UlyssesAttention(group, attn_fn)

    @functools.wraps(type(transformer).forward)
    def forward(self, x, t, text_states=None, text_mask=None, text_states_2=None, freqs_cos=None, freqs_sin=None,
                guidance=None, return_dict=True):
        P, r = sp._world(group), sp._rank(group)
        axis = next((a for a in (-2, -1) if (x.shape[a] // 2) % P == 0), None)  # 2 x 2 spatial patches
        if axis is None:
            raise ValueError(f"Cannot split video sequence into ulysses_degree x ring_degree ({P}) parts evenly")
        frames, rows, cols = x.shape[2], x.shape[3] // 2, x.shape[4] // 2

        def my_slab(table):
            if table is None:
                return None
            grid = table.reshape(frames, rows, cols, table.shape[-1])
            return grid.chunk(P, dim=axis - 1)[r].reshape(-1, table.shape[-1])

        for block in list(self.double_blocks) + list(self.single_blocks):
            block.hybrid_seq_parallel_attn = attn
        out = inner(x.chunk(P, dim=axis)[r], t, text_states, text_mask, text_states_2, my_slab(freqs_cos),
                    my_slab(freqs_sin), guidance, return_dict)
        if isinstance(out, dict):
            out["x"] = sp.gather_sequence(out["x"], dim=axis, group=group)
            return out
        return (sp.gather_sequence(out[0], dim=axis, group=group), *out[1:])

    transformer.forward = types.MethodType(forward, transformer)
    return transformer


def patch_sp(group=None) -> Dict[str, int]:
    """Make the reference's sequence-parallel entry points work without xfuser and on the b200vt kernels:
      * registers the xfuser stand-in (xfuser_shim.install(); no-op when the real package is importable), so that
        `from xfuser.core.distributed import ...` in xdit_context_parallel.py:3-7, flow/hunyuanvideo.py:31-46 and
        wan/text2video.py:262 resolves and `xFuserLongContextAttention()` is a b200vt.sp.UlyssesAttention;
      * rebinds `usp_attn_forward` / `usp_dit_forward` in wan/distributed/xdit_context_parallel.py (text2video.py:264-270
        imports them from there at bind time) to wan_usp_attn_forward / wan_usp_dit_forward;
      * rebinds `parallelize_transformer` in flow/hunyuanvideo.py and hyvideo_t2v/inference.py.
    Unsupported inputs fall back to the reference callables, which then run on the stand-in. Returns what was installed."""
    done = {"xfuser_shim": int(xfuser_shim.install()), "wan": 0, "hunyuan": 0}

    def _with_group(fast):
        if group is None:
            return fast
        return functools.wraps(fast)(lambda *a, **k: fast(*a, **{**k, "group": group}))

    mod = _try_import("videotuna.models.wan.wan.distributed.xdit_context_parallel")
    if mod is not None:
        mod.usp_attn_forward = _wrap(f"{mod.__name__}.usp_attn_forward", mod.usp_attn_forward,
                                     _with_group(wan_usp_attn_forward))
        mod.usp_dit_forward = _wrap(f"{mod.__name__}.usp_dit_forward", mod.usp_dit_forward, _with_group(wan_usp_dit_forward))
        done["wan"] = 2
    for name in ("videotuna.flow.hunyuanvideo", "videotuna.models.hunyuan.hyvideo_t2v.inference"):
        mod = _try_import(name)
        if mod is not None and hasattr(mod, "parallelize_transformer"):
            mod.parallelize_transformer = _wrap(f"{name}.parallelize_transformer", mod.parallelize_transformer,
                                                _with_group(hunyuan_parallelize_transformer))
            done["hunyuan"] += 1
    return done
