"""patch_videotuna(): hooks install on the importable reference tree, fall back to the untouched reference callable for
anything outside the CUDA path (here: CPU fp32 tensors), and restore cleanly. Runs only where /root/reference exists
(the development container); the GPU box has no reference tree."""
import os
import sys

import pytest
import torch

from oracle import ref_ops as R

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "videotuna")), reason="reference tree not present")


@pytest.fixture()
def shims():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    make_golden.install_shims()
    yield
    import b200vt.patch as P
    P.unpatch_videotuna()


@needs_ref
def test_lvdm_hook_falls_back_to_reference_on_cpu(shims):
    import b200vt.patch as P
    from videotuna.models.lvdm.modules import attention as A
    torch.manual_seed(0)
    m = A.CrossAttention(query_dim=128, context_dim=96, heads=2, dim_head=64)
    x, ctx = torch.randn(2, 50, 128), torch.randn(2, 80, 96)
    want = m(x, context=ctx)
    original = A.CrossAttention.forward
    assert P.patch_lvdm()
    assert A.CrossAttention.forward is not original and A.CrossAttention.forward._b200vt_original is original
    got = m(x, context=ctx)  # CPU fp32 -> Unsupported -> reference forward
    torch.testing.assert_close(got, want)
    assert P.patch_lvdm() and A.CrossAttention.forward._b200vt_original is original  # idempotent
    P.unpatch_videotuna()
    assert A.CrossAttention.forward is original


@needs_ref
def test_lvdm_xformers_switch_is_covered(shims):
    """The reference's spatial CrossAttention binds `self.forward = self.efficient_forward` when xformers is importable
    (attention.py:98-99). Instances built after patch_lvdm() must bind the replacement; instances built before it are
    fixed by rebind_lvdm_instances(); either way CPU fp32 inputs end in a reference body."""
    import b200vt.patch as P
    from videotuna.models.lvdm.modules import attention as A
    original_eff = A.CrossAttention.efficient_forward
    torch.manual_seed(0)
    early = A.CrossAttention(query_dim=128, context_dim=None, heads=2, dim_head=64)
    early.forward = early.efficient_forward  # what __init__ does with xformers installed
    assert P.patch_lvdm()
    assert A.CrossAttention.efficient_forward._b200vt_original is original_eff
    late = A.CrossAttention(query_dim=128, context_dim=None, heads=2, dim_head=64)
    late.forward = late.efficient_forward  # binds the patched class attribute
    assert getattr(late.forward.__func__, "_b200vt_patched", False)
    assert not getattr(early.forward.__func__, "_b200vt_patched", False)
    holder = torch.nn.Sequential(early)
    assert P.rebind_lvdm_instances(holder) == 1 and "forward" not in vars(early)
    x = torch.randn(2, 30, 128)
    want = A.CrossAttention.forward._b200vt_original(early, x)
    torch.testing.assert_close(early(x), want)  # class-level patched forward -> Unsupported on CPU -> reference forward
    P.unpatch_videotuna()
    assert A.CrossAttention.efficient_forward is original_eff


@needs_ref
def test_hunyuan_and_wan_hooks_install_and_fall_back(shims):
    import importlib
    import b200vt.patch as P
    done = P.patch_videotuna(lvdm=False)
    assert done["hunyuan"] >= 2  # models.py + attenion.py of at least one package
    mod = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.models")
    assert getattr(mod.attention, "_b200vt_patched", False)
    q, k, v = (torch.randn(1, 20, 2, 64) for _ in range(3))
    out = mod.attention(q, k, v, mode="torch")  # CPU -> reference SDPA path
    torch.testing.assert_close(out, R.hunyuan_attention_torch(q, k, v), rtol=1e-4, atol=1e-5)


def test_wan_sp_rope_tables_match_reference_slicing(monkeypatch):
    import b200vt.patch as P
    import b200vt.sp as sp
    freqs = R.wan_freqs_table(128)
    grid = (3, 4, 5)  # 60 tokens, padded to 64 = 2 ranks x 32
    full_cos, full_sin = R.wan_rope_cos_sin(grid, freqs)
    for rank in (0, 1):
        monkeypatch.setattr(sp, "_world", lambda g=None: 2)
        monkeypatch.setattr(sp, "_rank", lambda g=None, r=rank: r)
        cos, sin = P.wan_sp_rope_tables(grid, freqs, 32)
        lo, hi = rank * 32, min((rank + 1) * 32, 60)
        torch.testing.assert_close(cos[: hi - lo], full_cos[lo:hi])
        torch.testing.assert_close(sin[: hi - lo], full_sin[lo:hi])
        if hi - lo < 32:  # padding rows multiply by 1 + 0i (pad_freqs, xdit_context_parallel.py:12-22)
            assert torch.all(cos[hi - lo:] == 1) and torch.all(sin[hi - lo:] == 0)


@needs_ref
def test_block_hooks_install_fall_back_and_restore(shims):
    """patch_blocks(): every block-level forward is replaced on the reference classes; on CPU fp32 tensors the fused
    body raises Unsupported and the ORIGINAL forward runs, so results equal the unpatched reference bit for bit."""
    import importlib
    import b200vt.patch as P
    from videotuna.models.lvdm.modules import attention as A
    from videotuna.models.lvdm.modules.networks import openaimodel3d as O3
    M = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.models")
    torch.manual_seed(0)
    st = A.SpatialTransformer(in_channels=64, n_heads=1, d_head=64, depth=1, context_dim=32, use_linear=True,
                              use_checkpoint=False)
    x, ctx = torch.randn(2, 64, 4, 4), torch.randn(2, 7, 32)
    rb = O3.ResBlock(channels=32, emb_channels=16, dropout=0.0, out_channels=32, dims=2, use_checkpoint=False)
    xr, er = torch.randn(2, 32, 4, 4), torch.randn(2, 16)
    want_st, want_rb = st(x, ctx), rb(xr, er)
    originals = (A.SpatialTransformer.forward, A.BasicTransformerBlock._forward, O3.ResBlock._forward,
                 M.MMDoubleStreamBlock.forward)
    done = P.patch_blocks(wan=False)
    assert done["lvdm"] == 6 and done["hunyuan"] >= 3
    assert all(getattr(f, "_b200vt_patched", False) for f in (A.SpatialTransformer.forward, A.TemporalTransformer.forward,
                                                               A.BasicTransformerBlock._forward, O3.ResBlock._forward,
                                                               O3.TemporalConvBlock.forward,
                                                               M.MMDoubleStreamBlock.forward, M.MMSingleStreamBlock.forward,
                                                               M.parallel_attention))
    torch.testing.assert_close(st(x, ctx), want_st, rtol=0, atol=0)
    torch.testing.assert_close(rb(xr, er), want_rb, rtol=0, atol=0)
    P.unpatch_videotuna()
    assert (A.SpatialTransformer.forward, A.BasicTransformerBlock._forward, O3.ResBlock._forward,
            M.MMDoubleStreamBlock.forward) == originals


@needs_ref
def test_hunyuan_i2v_blocks_enter_the_fast_path_with_11_positionals(shims, monkeypatch):
    """HunyuanVideoFlow drives T2V and I2V through the i2v DiT, whose forward passes condition_type, token_replace_vec and
    frist_frame_token_num positionally to every block (hyvideo_i2v/modules/models.py:749-761, 776-788). The patched
    forwards must reach the drop-in bodies with that call (round 1 rejected it before the body ran); on CPU tensors the
    body then raises Unsupported and the original forward produces the reference result."""
    import importlib
    import b200vt.blocks as Bk
    import b200vt.functional as Fn
    import b200vt.patch as P
    M = importlib.import_module("videotuna.models.hunyuan.hyvideo_i2v.modules.models")
    entered = []

    def spy(name, real):
        def f(self, *a, **k):
            entered.append((name, len(a), a[8] if len(a) > 8 else None))
            return real(self, *a, **k)
        return f

    monkeypatch.setattr(Bk, "hunyuan_double_block_forward", spy("double", Bk.hunyuan_double_block_forward))
    monkeypatch.setattr(Bk, "hunyuan_single_block_forward", spy("single", Bk.hunyuan_single_block_forward))
    torch.manual_seed(0)
    dbl = M.MMDoubleStreamBlock(128, 2, mlp_width_ratio=1.0, qk_norm=True, qk_norm_type="rms", qkv_bias=True)
    sgl = M.MMSingleStreamBlock(128, 2, mlp_width_ratio=1.0, qk_norm=True, qk_norm_type="rms")
    img, txt, vec, trv = torch.randn(1, 40, 128), torch.randn(1, 8, 128), torch.randn(1, 128), torch.randn(1, 128)
    cu = torch.tensor([0, 48, 48], dtype=torch.int32)
    orig_attn = M.attention
    M.attention = lambda q_, k_, v_, **kw: orig_attn(q_, k_, v_, mode="torch")
    try:
        want = {c: (dbl(img, txt, vec, cu, cu, 48, 48, None, c, trv, 10),
                    sgl(torch.cat([img, txt], 1), vec, 8, cu, cu, 48, 48, None, c, trv, 10))
                for c in (None, "token_replace")}
        assert P.patch_blocks(lvdm=False, wan=False)["hunyuan"] >= 4
        for c in (None, "token_replace"):
            entered.clear()
            got_d = dbl(img, txt, vec, cu, cu, 48, 48, None, c, trv, 10)
            got_s = sgl(torch.cat([img, txt], 1), vec, 8, cu, cu, 48, 48, None, c, trv, 10)
            assert entered == [("double", 11, c), ("single", 11, c)]
            torch.testing.assert_close(got_d, want[c][0], rtol=0, atol=0)  # CPU -> Unsupported -> reference forward
            torch.testing.assert_close(got_s, want[c][1], rtol=0, atol=0)
        # the drop-in refuses what it cannot do BEFORE touching anything: token_replace without its vector
        with pytest.raises(Fn.Unsupported):
            Bk._token_replace("token_replace", None, 10)
        assert Bk._token_replace(None, trv, 10) == 0 and Bk._token_replace("token_replace", trv, 10) == 10
    finally:
        M.attention = orig_attn


def test_diffusers_processors_are_installed_by_duck_typing():
    """set_diffusers_processors() keys on the stock processor's class name and the `set_processor` protocol only
    (diffusers is not installed here); anything unsupported goes back to the stock processor."""
    import b200vt.patch as P

    class CogVideoXAttnProcessor2_0:  # stand-in for the stock processor
        def __call__(self, attn, hidden_states, encoder_hidden_states, attention_mask=None, image_rotary_emb=None):
            return "stock", None

    class Attention(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.processor = CogVideoXAttnProcessor2_0()
            self.heads = 2
            self.to_q = self.to_k = self.to_v = torch.nn.Linear(128, 128)

        def set_processor(self, p):
            self.processor = p

    model = torch.nn.Sequential(Attention(), Attention())
    assert P.set_diffusers_processors(model) == 2
    hs, ehs = torch.randn(1, 8, 128), torch.randn(1, 4, 128)
    assert model[0].processor(model[0], hs, ehs)[0] == "stock"  # CPU fp32 -> Unsupported -> stock processor


def test_diffusers_cogvideox_blocks_are_bound_per_instance_with_fallback():
    """set_diffusers_blocks() binds the drop-in forward on modules whose class is named CogVideoXBlock (diffusers is not
    installed here) and only once; CPU / fp32 activations raise Unsupported inside it and reach the stock forward."""
    import b200vt.patch as P

    class CogVideoXBlock(torch.nn.Module):  # stand-in: only the name and a forward matter for the binding
        def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb=None, attention_kwargs=None):
            return "stock", hidden_states.shape

    model = torch.nn.Sequential(CogVideoXBlock(), torch.nn.Linear(2, 2), CogVideoXBlock())
    assert P.set_diffusers_blocks(model) == 2
    assert P.set_diffusers_blocks(model) == 0  # idempotent
    hs, ehs, temb = torch.randn(1, 8, 128), torch.randn(1, 4, 128), torch.randn(1, 16)
    assert model[0](hs, ehs, temb)[0] == "stock"
    assert model[2](hs, ehs, temb=temb)[0] == "stock"


# ---- LoRA projections as dense GEMMs (functional.lora_merged_projections) ------------------------------------------------
def test_lora_projections_as_dense_gemms_match_the_adapter_forward():
    import b200vt.functional as Fn
    from helpers import PeftLikeLinear as _PeftLikeLinear
    torch.manual_seed(3)
    lin = lambda bias: _PeftLikeLinear(torch.nn.Linear(48, 32, bias=bias))  # noqa: E731
    for bias in (False, True):
        layers = [lin(bias) for _ in range(3)]
        x = torch.randn(2, 7, 48, requires_grad=True)
        want = [l(x) for l in layers]
        gy = [torch.randn_like(w) for w in want]
        torch.autograd.backward(want, gy)
        ref = {"x": x.grad.clone(), **{f"{i}{n}": p.grad.clone() for i, l in enumerate(layers) for n, p in l.named_parameters() if p.grad is not None}}
        x.grad = None
        for l in layers:
            l.zero_grad()
        got = Fn.lora_merged_projections(x, layers)
        assert got is not None and len(got) == 3
        for g, w in zip(got, want):
            torch.testing.assert_close(g, w, rtol=1e-5, atol=1e-5)
        torch.autograd.backward(got, gy)
        torch.testing.assert_close(x.grad, ref["x"], rtol=1e-4, atol=1e-5)
        for i, l in enumerate(layers):
            for n, p in l.named_parameters():
                if f"{i}{n}" in ref:
                    torch.testing.assert_close(p.grad, ref[f"{i}{n}"], rtol=1e-4, atol=1e-5)
                else:
                    assert p.grad is None  # frozen base weights stay without gradient
    # not eligible: plain linears only (nothing to merge), dropout, merged / disabled adapters, trainable base
    assert Fn.lora_merged_projections(x, [torch.nn.Linear(48, 32)]) is None
    assert Fn.lora_merged_projections(x, [_PeftLikeLinear(torch.nn.Linear(48, 32), dropout=0.1)]) is None
    m = lin(False)
    m.merged = True
    assert Fn.lora_merged_projections(x, [m]) is None
    m = lin(False)
    m.base_layer.weight.requires_grad_(True)
    assert Fn.lora_merged_projections(x, [m]) is None
    # a plain nn.Linear next to adapters is carried with a zero delta
    mixed = Fn.lora_merged_projections(x, [lin(False), torch.nn.Linear(48, 32, bias=False).requires_grad_(False)])
    assert mixed is not None and mixed[1].shape == (2, 7, 32)


def test_cast_frozen_weights_touches_only_frozen_conv_and_linear():
    import b200vt.patch as P
    net = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.GroupNorm(2, 8), torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    for p in net.parameters():
        p.requires_grad_(False)
    net[3].weight.requires_grad_(True)  # a trainable tensor (e.g. an adapter) keeps fp32
    want = net[0].weight.detach().clone().to(torch.bfloat16)
    n = P.cast_frozen_weights(net)
    assert n == 5  # conv w+b, linear w+b, last linear's bias
    assert net[0].weight.dtype == torch.bfloat16 and torch.equal(net[0].weight, want)  # the rounding autocast applies
    assert net[1].weight.dtype == torch.float32 and net[3].weight.dtype == torch.float32 and net[3].bias.dtype == torch.bfloat16


def test_ckpt_policy_and_checkpoint_wrapper():
    """b200vt.ckpt on CPU: the policy keeps exactly the attention forward ops, ckpt.checkpoint runs torch's non-reentrant
    checkpoint with that policy, and keep_attention_in_checkpoints() wraps / restores torch.utils.checkpoint.checkpoint."""
    import torch.utils.checkpoint as tuc
    import b200vt.ckpt as CK
    import b200vt.sp  # noqa: F401  (registers b200vt::ulysses_attn_fwd)
    keep, redo = tuc.CheckpointPolicy.MUST_SAVE, tuc.CheckpointPolicy.PREFER_RECOMPUTE
    assert CK.attention_saving_policy(None, torch.ops.b200vt.attn_fwd.default) == keep
    assert CK.attention_saving_policy(None, torch.ops.b200vt.ulysses_attn_fwd.default) == keep
    assert CK.attention_saving_policy(None, torch.ops.b200vt.attn_bwd.default) == redo
    assert CK.attention_saving_policy(None, torch.ops.aten.mm.default) == redo
    lin = torch.nn.Linear(8, 8)
    x = torch.randn(2, 8, requires_grad=True)
    want = torch.autograd.grad(lin(x).sum(), x)[0]
    got = torch.autograd.grad(CK.checkpoint(lin, x).sum(), x)[0]
    torch.testing.assert_close(got, want)
    with pytest.raises(ValueError):
        CK.checkpoint(lin, x, use_reentrant=True)
    CK.keep_attention_in_checkpoints()
    try:
        assert getattr(tuc.checkpoint, "_b200vt_wrapped", False)
        for reentrant in (False, True):
            x.grad = None
            torch.utils.checkpoint.checkpoint(lin, x, use_reentrant=reentrant).sum().backward()
            torch.testing.assert_close(x.grad, want)
    finally:
        CK.keep_attention_in_checkpoints(False)
    assert tuc.checkpoint is CK._ORIGINAL


def test_groupnorm_ops_bind_defaulted_arguments_under_fake_tensors():
    """The GroupNorm ops gained defaulted arguments (`addend`, `need_wgrad`): traced / fake-tensor calls with the old six
    positional arguments must still bind (torch fills the defaults before setup_context), the layout of the fake output
    must follow the input's (channels-last stays channels-last), and the backward formula must return one entry per input."""
    import types
    from torch._subclasses.fake_tensor import FakeTensorMode
    import b200vt.ops as ops
    with FakeTensorMode():
        for cl, use_e, frozen, nargs in ((False, False, False, 6), (True, False, False, 6), (True, True, False, 7), (True, True, True, 7)):
            x = torch.empty(3, 64, 6, 10, device="cuda", dtype=torch.bfloat16)
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
            w, b = torch.empty(64, device="cuda"), torch.empty(64, device="cuda")
            e = torch.empty(3, 64, device="cuda") if use_e else None
            args = (x, w, b, 32, 1e-5, True) + ((e,) if nargs == 7 else ())
            y, mean, rstd = ops.groupnorm_silu_fwd(*args)
            assert y.shape == x.shape and y.stride() == x.stride() and mean.shape == (3, 32)
            ctx = types.SimpleNamespace(saved_tensors=(x, mean, rstd, w, b, e), groups=32, silu=True,
                                        needs_input_grad=(True, not frozen, not frozen, False, False, False, False))
            out = ops._gn_backward(ctx, torch.empty_like(y), None, None)
            assert len(out) == 7 and out[0].shape == x.shape and out[0].stride() == x.stride()
            assert (out[1] is None and out[2] is None) if frozen else (out[1].shape == (64,) and out[2].shape == (64,))
