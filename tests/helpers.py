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

    @classmethod
    def from_fixture(cls, case, device, dtype=torch.bfloat16):
        m = cls(**case["kw"])
        sd = {k: v for k, v in case["sd"].items() if not k.startswith("relative_position")}
        m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
        return m.to(device=device, dtype=dtype)
