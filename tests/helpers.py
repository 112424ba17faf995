"""Stand-ins for reference modules, so GPU tests (which cannot import /root/reference) can drive the drop-in bodies.

`CrossAttentionShell` carries exactly the attributes and parameter names lvdm's CrossAttention.__init__ creates
(videotuna/models/lvdm/modules/attention.py:45-99) — and nothing of its forward: the fixtures in tests/golden hold the
reference's outputs for the same state dict."""
import torch
from torch import nn


class CrossAttentionShell(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0, relative_position=False,
                 temporal_length=None, img_cross_attention=False, img_cross_attention_scale=1.0,
                 img_cross_attention_scale_learnable=False, text_context_len=77):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))
        self.img_cross_attention = img_cross_attention
        self.img_cross_attention_scale = img_cross_attention_scale
        self.img_cross_attention_scale_learnable = img_cross_attention_scale_learnable
        self.text_context_len = text_context_len
        if img_cross_attention:
            self.to_k_ip = nn.Linear(context_dim, inner, bias=False)
            self.to_v_ip = nn.Linear(context_dim, inner, bias=False)
            if img_cross_attention_scale_learnable:
                self.register_parameter("alpha", nn.Parameter(torch.tensor(0.0)))
        self.relative_position = relative_position
        if relative_position:
            self.relative_position_k = RelativePositionShell(dim_head, temporal_length)
            self.relative_position_v = RelativePositionShell(dim_head, temporal_length)

    @classmethod
    def from_fixture(cls, case, device, dtype=torch.bfloat16):
        m = cls(**case["kw"])
        m.load_state_dict({k: v.float() for k, v in case["sd"].items()}, strict=True)
        return m.to(device=device, dtype=dtype)


class RelativePositionShell(nn.Module):
    """lvdm RelativePosition (attention.py:19-42): a (2 * max + 1, D) table gathered at clamp(k - q, -max, max) + max."""

    def __init__(self, num_units, max_relative_position):
        super().__init__()
        self.num_units, self.max_relative_position = num_units, max_relative_position
        self.embeddings_table = nn.Parameter(torch.zeros(max_relative_position * 2 + 1, num_units))

    def forward(self, length_q, length_k):
        dev = self.embeddings_table.device
        dist = torch.arange(length_k, device=dev)[None, :] - torch.arange(length_q, device=dev)[:, None]
        idx = dist.clamp(-self.max_relative_position, self.max_relative_position) + self.max_relative_position
        return self.embeddings_table[idx.long()]


# ---------------------------------------------------------------------------------------------------------------------
# Shells of the reference BLOCKS: same attribute and parameter names as the reference constructors create, with the
# drop-in forwards of b200vt.blocks bound the way patch.patch_blocks() binds them onto the real classes.
# ---------------------------------------------------------------------------------------------------------------------
def _load(module, sd, device, dtype=torch.bfloat16):
    module.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}, strict=True)
    return module.to(device=device, dtype=dtype)


class RMSNormShell(nn.Module):
    """hunyuan norm_layers.RMSNorm / wan WanRMSNorm: .weight, .eps (the forward is never called by the drop-ins)."""

    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class ModulateDiTShell(nn.Module):
    def __init__(self, hidden, factor):
        super().__init__()
        self.act = nn.SiLU()
        self.linear = nn.Linear(hidden, factor * hidden)

    def forward(self, x):
        return self.linear(self.act(x))


class MLPShell(nn.Module):
    def __init__(self, hidden, mlp_hidden):
        super().__init__()
        self.fc1 = nn.Linear(hidden, mlp_hidden)
        self.act = nn.GELU(approximate="tanh")
        self.fc2 = nn.Linear(mlp_hidden, hidden)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class HunyuanDoubleShell(nn.Module):
    """MMDoubleStreamBlock.__init__ (hyvideo_t2v/modules/models.py:28-128)."""

    def __init__(self, hidden, heads, mlp_width_ratio):
        super().__init__()
        from b200vt import blocks
        self.heads_num = heads
        d = hidden // heads
        for s in ("img", "txt"):
            setattr(self, f"{s}_mod", ModulateDiTShell(hidden, 6))
            setattr(self, f"{s}_norm1", nn.LayerNorm(hidden, elementwise_affine=False, eps=1e-6))
            setattr(self, f"{s}_attn_qkv", nn.Linear(hidden, 3 * hidden))
            setattr(self, f"{s}_attn_q_norm", RMSNormShell(d))
            setattr(self, f"{s}_attn_k_norm", RMSNormShell(d))
            setattr(self, f"{s}_attn_proj", nn.Linear(hidden, hidden))
            setattr(self, f"{s}_norm2", nn.LayerNorm(hidden, elementwise_affine=False, eps=1e-6))
            setattr(self, f"{s}_mlp", MLPShell(hidden, int(hidden * mlp_width_ratio)))
        self.hybrid_seq_parallel_attn = None
        self._fwd = blocks.hunyuan_double_block_forward

    def forward(self, *a, **k):
        return self._fwd(self, *a, **k)


class HunyuanSingleShell(nn.Module):
    """MMSingleStreamBlock.__init__ (models.py:262-318)."""

    def __init__(self, hidden, heads, mlp_width_ratio):
        super().__init__()
        from b200vt import blocks
        self.hidden_size, self.heads_num = hidden, heads
        self.mlp_hidden_dim = int(hidden * mlp_width_ratio)
        d = hidden // heads
        self.scale = d ** -0.5
        self.linear1 = nn.Linear(hidden, 3 * hidden + self.mlp_hidden_dim)
        self.linear2 = nn.Linear(hidden + self.mlp_hidden_dim, hidden)
        self.q_norm, self.k_norm = RMSNormShell(d), RMSNormShell(d)
        self.pre_norm = nn.LayerNorm(hidden, elementwise_affine=False, eps=1e-6)
        self.mlp_act = nn.GELU(approximate="tanh")
        self.modulation = ModulateDiTShell(hidden, 3)
        self.hybrid_seq_parallel_attn = None
        self._fwd = blocks.hunyuan_single_block_forward

    def forward(self, *a, **k):
        return self._fwd(self, *a, **k)


class WanAttnShell(nn.Module):
    """WanSelfAttention / WanT2VCrossAttention.__init__ (wan/wan/modules/model.py:102-124)."""

    def __init__(self, dim, heads, cross, eps=1e-6):
        super().__init__()
        from b200vt import blocks
        self.dim, self.num_heads, self.head_dim = dim, heads, dim // heads
        self.window_size, self.qk_norm, self.eps = (-1, -1), True, eps
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = RMSNormShell(dim, eps), RMSNormShell(dim, eps)
        self._fwd = blocks.wan_t2v_cross_attention_forward if cross else blocks.wan_self_attention_forward

    def forward(self, *a, **k):
        return self._fwd(self, *a, **k)


class WanBlockShell(nn.Module):
    """WanAttentionBlock.__init__ (model.py:230-272), t2v cross-attention, cross_attn_norm=True."""

    def __init__(self, dim, ffn, heads, eps=1e-6):
        super().__init__()
        from b200vt import blocks
        self.dim, self.ffn_dim, self.num_heads, self.eps = dim, ffn, heads, eps
        self.norm1 = nn.LayerNorm(dim, eps, elementwise_affine=False)
        self.self_attn = WanAttnShell(dim, heads, cross=False, eps=eps)
        self.norm3 = nn.LayerNorm(dim, eps, elementwise_affine=True)
        self.cross_attn = WanAttnShell(dim, heads, cross=True, eps=eps)
        self.norm2 = nn.LayerNorm(dim, eps, elementwise_affine=False)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn), nn.GELU(approximate="tanh"), nn.Linear(ffn, dim))
        self.modulation = nn.Parameter(torch.zeros(1, 6, dim))
        self._fwd = blocks.wan_attention_block_forward

    def forward(self, *a, **k):
        return self._fwd(self, *a, **k)


class GEGLUShell(nn.Module):
    """lvdm GEGLU (attention.py:522-529) with the drop-in forward patch_blocks() installs (functional.lvdm_geglu_forward);
    inputs outside the CUDA path take the reference arithmetic, as the patch layer's fallback does."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        from b200vt import functional
        try:
            return functional.lvdm_geglu_forward(self, x)
        except functional.Unsupported:
            x, gate = self.proj(x).chunk(2, dim=-1)
            return x * torch.nn.functional.gelu(gate)


class FeedForwardShell(nn.Module):
    """lvdm FeedForward(dim, glu=True) (attention.py:222-242): net = [GEGLU, Dropout, Linear]."""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(GEGLUShell(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))

    def forward(self, x):
        return self.net(x)


class BasicBlockShell(nn.Module):
    """lvdm BasicTransformerBlock.__init__ (attention.py:245-281); forward reproduces the reference's argument plumbing
    (:283-297, incl. dropping the context when a mask is given) without the checkpoint wrapper."""

    def __init__(self, dim, n_heads, d_head, context_dim=None):
        super().__init__()
        from b200vt import blocks, functional
        self.disable_self_attn = False
        self.attn1 = CrossAttentionShell(dim, None, n_heads, d_head)
        self.ff = FeedForwardShell(dim)
        self.attn2 = CrossAttentionShell(dim, context_dim, n_heads, d_head)
        for a in (self.attn1, self.attn2):
            a.forward = functional.lvdm_cross_attention_forward.__get__(a)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.checkpoint = False
        self._fwd = blocks.lvdm_basic_block_forward

    def forward(self, x, context=None, mask=None):
        if self.checkpoint and torch.is_grad_enabled():  # the reference checkpoints here (attention.py:283-297, utils.py:112-125)
            from torch.utils.checkpoint import checkpoint
            if mask is not None:
                return checkpoint(lambda t: self._fwd(self, t, mask=mask), x, use_reentrant=False)
            if context is not None:
                return checkpoint(lambda t, c: self._fwd(self, t, c), x, context, use_reentrant=False)
            return checkpoint(lambda t: self._fwd(self, t), x, use_reentrant=False)
        if mask is not None:
            return self._fwd(self, x, mask=mask)
        return self._fwd(self, x, context) if context is not None else self._fwd(self, x)


class SpatialTransformerShell(nn.Module):
    """lvdm SpatialTransformer.__init__ (attention.py:323-374), use_linear=True."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, context_dim=None, use_linear=True, use_checkpoint=False):
        super().__init__()
        from b200vt import blocks
        inner = n_heads * d_head
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(BasicBlockShell(inner, n_heads, d_head, context_dim) for _ in range(depth))
        self.proj_out = nn.Linear(inner, in_channels)
        self.use_linear = use_linear
        self._fwd = blocks.lvdm_spatial_transformer_forward

    def forward(self, *a, **k):
        return self._fwd(self, *a, **k)


class TemporalTransformerShell(nn.Module):
    """lvdm TemporalTransformer.__init__ (attention.py:403-473), only_self_att=True; use_linear=False gives the Conv1d
    projections the UNet's init_attn is built with (openaimodel3d.py:418-432 passes no use_linear)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, use_linear=True, use_checkpoint=False, only_self_att=True,
                 temporal_length=None, causal_attention=False):
        super().__init__()
        from b200vt import blocks
        inner = n_heads * d_head
        self.only_self_att, self.causal_attention, self.relative_position = only_self_att, causal_attention, False
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear else nn.Conv1d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(BasicBlockShell(inner, n_heads, d_head, None) for _ in range(depth))
        self.proj_out = nn.Linear(inner, in_channels) if use_linear else nn.Conv1d(inner, in_channels, 1)
        self.use_linear = use_linear
        if causal_attention:
            self.mask = torch.tril(torch.ones([1, temporal_length, temporal_length]))
        self._fwd = blocks.lvdm_temporal_transformer_forward

    def forward(self, *a, **k):
        return self._fwd(self, *a, **k)


class ResBlockShell(nn.Module):
    """lvdm ResBlock.__init__ (openaimodel3d.py:139-210), dims=2, no up/down-sampling; use_temporal_conv adds the
    TemporalConvBlock under the reference's attribute name (`temopral_conv`, :205-210, dropout 0.1)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, dims=2, use_checkpoint=False,
                 use_temporal_conv=False):
        super().__init__()
        from b200vt import blocks
        self.channels, self.emb_channels, self.out_channels = channels, emb_channels, out_channels or channels
        self.use_checkpoint, self.use_scale_shift_norm, self.use_temporal_conv = use_checkpoint, False, use_temporal_conv
        self.updown = False
        self.in_layers = nn.Sequential(nn.GroupNorm(32, channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(nn.GroupNorm(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1))
        self.skip_connection = (nn.Identity() if self.out_channels == channels
                                else nn.Conv2d(channels, self.out_channels, 1))
        if use_temporal_conv:
            self.temopral_conv = TemporalConvBlockShell(self.out_channels, dropout=0.1)
        self._fwd = blocks.lvdm_resblock_forward

    def forward(self, x, emb, batch_size=None):
        return self._fwd(self, x, emb, batch_size)


class TemporalConvBlockShell(nn.Module):
    """lvdm TemporalConvBlock.__init__ (openaimodel3d.py:263-301), spatial_aware=False."""

    def __init__(self, channels, dropout=0.0):
        super().__init__()
        from b200vt import blocks

        def stage(with_dropout):
            layers = [nn.GroupNorm(32, channels), nn.SiLU()] + ([nn.Dropout(dropout)] if with_dropout else [])
            return nn.Sequential(*layers, nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0)))

        self.conv1, self.conv2, self.conv3, self.conv4 = stage(False), stage(True), stage(True), stage(True)
        self._fwd = blocks.lvdm_temporal_conv_block_forward

    def forward(self, x):
        return self._fwd(self, x)


class DiffusersAttentionShell(nn.Module):
    """diffusers 0.32.2 `Attention` as CogVideoXBlock builds it (query_dim=dim, heads, dim_head, qk_norm="layer_norm",
    eps=1e-6, bias=True, out_bias=True): to_q/to_k/to_v, to_out = [Linear, Dropout], norm_q/norm_k = LayerNorm(dim_head),
    .heads, .processor, set_processor(); forward hands over to the processor like the real module."""

    def __init__(self, dim, heads, processor=None):
        super().__init__()
        self.heads = heads
        d = dim // heads
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.norm_q, self.norm_k = nn.LayerNorm(d, eps=1e-6), nn.LayerNorm(d, eps=1e-6)
        self.is_cross_attention = False
        self.processor = processor

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class CogLayerNormZeroShell(nn.Module):
    """diffusers `CogVideoXLayerNormZero(conditioning_dim, embedding_dim, elementwise_affine=True, eps=1e-5, bias=True)`."""

    def __init__(self, cond_dim, dim, eps=1e-5):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(cond_dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=True)


class _GELUProj(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return torch.nn.functional.gelu(self.proj(x), approximate="tanh")


class DiffusersFeedForwardShell(nn.Module):
    """diffusers `FeedForward(dim, activation_fn="gelu-approximate")`: net = [GELU(proj), Dropout, Linear]."""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class CogVideoXBlockShell(nn.Module):
    """diffusers 0.32.2 `CogVideoXBlock.__init__` (dim, num_attention_heads, attention_head_dim, time_embed_dim): norm1,
    attn1, norm2, ff, with the drop-in forward bound as patch.set_diffusers_blocks() binds it."""

    def __init__(self, dim, heads, time_embed_dim, processor=None):
        super().__init__()
        from b200vt import blocks
        self.norm1 = CogLayerNormZeroShell(time_embed_dim, dim)
        self.attn1 = DiffusersAttentionShell(dim, heads, processor if processor is not None else blocks.CogVideoXAttnProcessor())
        self.norm2 = CogLayerNormZeroShell(time_embed_dim, dim)
        self.ff = DiffusersFeedForwardShell(dim)
        self._fwd = blocks.cogvideox_block_forward

    def forward(self, *a, **k):
        return self._fwd(self, *a, **k)


def load_shell(module, sd, device, dtype=torch.bfloat16):
    return _load(module, sd, device, dtype)


class PeftLikeLinear(torch.nn.Module):
    """The attributes and forward of peft.tuners.lora.Linear that matter here (peft is not installed in this image):
    ModuleDict adapters keyed by name, per-adapter scaling / dropout, `base(x) + lora_B(lora_A(dropout(x))) * scaling`."""

    def __init__(self, base, r=4, alpha=1.0, dropout=0.0, name="default"):
        super().__init__()
        self.base_layer = base
        base.weight.requires_grad_(False)
        if base.bias is not None:
            base.bias.requires_grad_(False)
        self.lora_A = torch.nn.ModuleDict({name: torch.nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = torch.nn.ModuleDict({name: torch.nn.Linear(r, base.out_features, bias=False)})
        self.lora_dropout = torch.nn.ModuleDict({name: torch.nn.Dropout(dropout) if dropout > 0 else torch.nn.Identity()})
        self.scaling = {name: alpha / r}
        self.active_adapters = [name]
        self.use_dora = {name: False}
        self.merged = False
        self.disable_adapters = False
        torch.nn.init.normal_(self.lora_B[name].weight, std=0.05)

    def forward(self, x):
        n = self.active_adapters[0]
        return self.base_layer(x) + self.lora_B[n](self.lora_A[n](self.lora_dropout[n](x))) * self.scaling[n]
