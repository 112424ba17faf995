"""The benchmark tools compare the b200vt path with "the reference's own op sequence" written out inside the tool
(tools/bench_vc2_blocks.py, tools/bench_denoiser.py). These tests pin those restatements to the unmodified reference
classes on CPU, so the comparison arm is the reference's arithmetic and not an approximation of it. Development container
only (needs /root/reference)."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "videotuna")), reason="reference tree not present")


def _load_tool(name):
    argv, sys.argv = sys.argv, ["x"]
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.argv = argv


def _dezero(m):
    for p in m.parameters():
        if float(p.detach().abs().sum()) == 0.0:
            torch.nn.init.normal_(p, std=0.02)
    return m


@pytest.fixture()
def ref_env():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    make_golden.install_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    yield


@needs_ref
def test_vc2_blocks_torch_arm_equals_reference_modules(ref_env):
    bv = _load_tool("bench_vc2_blocks")
    from videotuna.models.lvdm.modules import attention as A
    from videotuna.models.lvdm.modules.networks import openaimodel3d as O
    torch.manual_seed(0)
    ref = _dezero(A.SpatialTransformer(64, 2, 32, depth=1, context_dim=48, use_linear=True, use_checkpoint=False))
    sh = bv.H.SpatialTransformerShell(64, 2, 32, depth=1, context_dim=48)
    sh.load_state_dict(ref.state_dict(), strict=True)
    x, ctx = torch.randn(2, 64, 6, 5), torch.randn(2, 80, 48)  # 80 > 77: the text-context truncation is exercised
    torch.testing.assert_close(bv.spatial_torch(sh, x, ctx), ref(x, ctx), rtol=1e-5, atol=1e-6)

    ref = _dezero(A.TemporalTransformer(64, 2, 32, depth=1, use_linear=True, use_checkpoint=False, only_self_att=True,
                                        temporal_length=4))
    sh = bv.H.TemporalTransformerShell(64, 2, 32, depth=1, temporal_length=4)
    sh.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 64, 4, 3, 5)
    torch.testing.assert_close(bv.temporal_torch(sh, x), ref(x), rtol=1e-5, atol=1e-6)

    ref = _dezero(O.ResBlock(64, 32, 0.0, out_channels=64, dims=2, use_checkpoint=False, use_temporal_conv=False))
    sh = bv.H.ResBlockShell(64, 32, 0.0)
    sh.load_state_dict(ref.state_dict(), strict=True)
    x, emb = torch.randn(3, 64, 6, 5), torch.randn(3, 32)
    torch.testing.assert_close(bv.resblock_torch(sh, x, emb), ref(x, emb), rtol=1e-5, atol=1e-6)


def _varlen_sdpa(q, k, v, cu_q, cu_k, max_q, max_k):
    """CPU stand-in for flash_attn_varlen_func: packed (total, H, D) tensors, attention inside each [cu[i], cu[i+1]) segment."""
    out = torch.zeros_like(q)
    for i in range(len(cu_q) - 1):
        a, b, c, d = int(cu_q[i]), int(cu_q[i + 1]), int(cu_k[i]), int(cu_k[i + 1])
        if b > a and d > c:
            o = torch.nn.functional.scaled_dot_product_attention(q[a:b].transpose(0, 1)[None].float(),
                                                                 k[c:d].transpose(0, 1)[None].float(),
                                                                 v[c:d].transpose(0, 1)[None].float())
            out[a:b] = o[0].transpose(0, 1).to(q.dtype)
    return out


def test_denoiser_torch_arms_equal_reference_block_fixtures():
    """bench_denoiser's `torch` arm (Hunyuan double / single stream block, Wan attention block) on the state dicts and
    inputs of the committed fixtures, which hold the outputs of the unmodified reference blocks (tests/golden)."""
    from conftest import load_golden
    from helpers import HunyuanDoubleShell, HunyuanSingleShell, WanBlockShell
    from oracle import ref_ops as R
    bd = _load_tool("bench_denoiser")
    bd._flash_varlen = _varlen_sdpa
    g = load_golden("hunyuan_blocks")
    hidden, heads = g["hidden"], g["heads"]
    dbl = HunyuanDoubleShell(hidden, heads, g["mlp_width_ratio"]).float()
    dbl.load_state_dict({k: v.float() for k, v in g["dbl_sd"].items()}, strict=True)
    sgl = HunyuanSingleShell(hidden, heads, g["mlp_width_ratio"]).float()
    sgl.load_state_dict({k: v.float() for k, v in g["sgl_sd"].items()}, strict=True)
    img, txt, vec, cu = g["img"].float(), g["txt"].float(), g["vec"].float(), g["cu_seqlens"]
    fc = (g["cos"].float(), g["sin"].float())
    S = img.shape[1] + txt.shape[1]
    io, to = bd.hy_double_torch(dbl, img, txt, vec, cu, cu, S, S, fc)
    assert R.max_rel_err(io, g["img_out"].float()) < 2e-3 and R.max_rel_err(to, g["txt_out"].float()) < 2e-3
    so = bd.hy_single_torch(sgl, torch.cat([img, txt], 1), vec, txt.shape[1], cu, cu, S, S, fc)
    assert R.max_rel_err(so, g["single_out"].float()) < 2e-3

    w = load_golden("wan_block")
    blk = WanBlockShell(w["dim"], w["ffn"], w["heads"]).float()
    blk.load_state_dict({k: v.float() for k, v in w["sd"].items()}, strict=True)
    x = w["x"].float()
    freqs = bd.wan_freqs_table(w["dim"] // w["heads"], "cpu")
    y = bd.wan_block_torch(blk, x, w["e"].float(), torch.tensor([x.shape[1]]), w["grid"], freqs, w["context"].float(), None)
    assert R.max_rel_err(y, w["out"].float()) < 2e-3


@needs_ref
def test_vc2_unet_shell_and_torch_arm_equal_the_reference_unet(ref_env):
    """tools/bench_vc2_unet.py builds the VideoCrafter2 3D-UNet from shells (the GPU box has no reference tree) and drives it
    with the reference's op sequence in its `torch` arm: the UNMODIFIED UNetModel's state dict must load strictly into the
    shell, and the torch arm must reproduce the reference forward (a reduced configuration of the same family, fp32, CPU)."""
    import importlib.util
    from videotuna.models.lvdm.modules.networks.openaimodel3d import UNetModel
    spec = importlib.util.spec_from_file_location("bench_vc2_unet", os.path.join(ROOT, "tools", "bench_vc2_unet.py"))
    U = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(U)
    cfg = dict(in_channels=4, out_channels=4, model_channels=64, attention_resolutions=(2, 1), num_res_blocks=1,
               channel_mult=(1, 2), num_head_channels=32, transformer_depth=1, context_dim=48, temporal_length=4,
               temporal_conv=True, addition_attention=True, fps_cond=True)
    torch.manual_seed(0)
    ref = UNetModel(use_linear=True, use_checkpoint=False, temporal_attention=True, temporal_selfatt_only=True,
                    use_relative_position=False, use_causal_attention=False, **cfg).eval()
    with torch.no_grad():
        for p in ref.parameters():  # zero-initialised output layers re-drawn: a vacuous comparison otherwise
            if float(p.abs().max()) == 0.0:
                p.copy_(torch.randn_like(p) * 0.05)
    shell = U.VC2UNet(**cfg).eval()
    shell.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 4, 8, 8, generator=g)
    ctx = torch.randn(2, 7, 48, generator=g)
    t = torch.tensor([17, 503])
    with torch.no_grad():
        want = ref(x, t, context=ctx, fps=24)
        got = shell(x, t, ctx, 24, ours=False, ckpt=False)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
