"""The reference's sequence-parallel ENTRY POINTS running without xfuser, on CPU: world_size-2 gloo processes.

    wan      usp_dit_forward / usp_attn_forward   videotuna/models/wan/wan/distributed/xdit_context_parallel.py:66-192
    hunyuan  parallelize_transformer              videotuna/flow/hunyuanvideo.py:114-178 (needs the whole flow to import; its
                                                  drop-in b200vt.patch.hunyuan_parallelize_transformer is tested on the real
                                                  hyvideo_i2v DiT instead), parallel_attention attenion.py:159-212
Both the b200vt drop-ins (patch.patch_sp()) and the UNMODIFIED reference wrappers — imported on top of the xfuser stand-in
(b200vt.xfuser_shim) — must reproduce the single-process reference forward of the same model on the same inputs: SP is a
pure re-partitioning. The attention core is injected (oracle softmax attention) because the CUDA kernels
cannot run here; what is under test is the host logic: stand-in group bookkeeping, token chunking, per-rank RoPE tables,
replicated text ("rear"), padding tail, final gather. Needs /root/reference (development container only)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_ops as R

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "videotuna")), reason="reference tree not present")
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_attn(q, k, v, softmax_scale=None):
    # fp32 result whatever the input dtype: the reference's usp_attn_forward hands bf16 q, k, v to the attention object and
    # feeds the result to an fp32 Linear in this CPU test (bf16 weights / autocast in the real pipeline)
    return R.sdpa_blhd(q.float(), k.float(), v.float(), None, softmax_scale)


def _shims():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    make_golden.install_shims()


# ---------------------------------------------------------------------------------------------------------------------
# Wan
# ---------------------------------------------------------------------------------------------------------------------
def _wan_modules():
    import importlib
    import types
    base = os.path.join(REF, "videotuna/models/wan/wan")
    for name, sub in (("wan_ref", ""), ("wan_ref.modules", "modules"), ("wan_ref.distributed", "distributed")):
        pkg = types.ModuleType(name)  # synthetic packages: skip wan/__init__.py and modules/__init__.py (easydict, T5, ftfy)
        pkg.__path__ = [os.path.join(base, sub)]
        sys.modules[name] = pkg
    model = importlib.import_module("wan_ref.modules.model")

    def sdpa_flash(q, k, v, q_lens=None, k_lens=None, dropout_p=0.0, softmax_scale=None, q_scale=None, causal=False,
                   window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
        # the reference's own SDPA fallback body (wan/modules/attention.py:171-179); flash_attention asserts CUDA. Like
        # flash_attention (:59-83) it rounds q, k, v to bf16 first and returns q's dtype.
        q, k, v = (t.to(dtype).float() for t in (q, k, v))
        o = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2).contiguous()

    model.flash_attention = sdpa_flash
    return model


def _wan_model(model):
    torch.manual_seed(7)
    m = model.WanModel(model_type="t2v", patch_size=(1, 2, 2), text_len=12, in_dim=4, dim=128, ffn_dim=256, freq_dim=64,
                       text_dim=32, out_dim=4, num_heads=2, num_layers=2, qk_norm=True, cross_attn_norm=True)
    with torch.no_grad():  # head / modulation are zero- or tiny-initialised: make the parity non-vacuous
        for p in m.parameters():
            if float(p.abs().max()) < 1e-3:
                p.copy_(torch.randn_like(p) * 0.05)
    return m.eval()


def _wan_inputs():
    g = torch.Generator().manual_seed(11)
    x = [torch.randn(4, 3, 8, 8, generator=g)]                 # (C, F, H, W): 3 * 4 * 4 = 48 tokens
    ctx = [torch.randn(9, 32, generator=g)]
    return x, torch.tensor([417.0]), ctx, 48


def _wan_worker(rank, world, port, variant, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        import types
        import b200vt.patch as P
        import b200vt.xfuser_shim as X
        _shims()
        model = _wan_modules()
        X.ATTN_FN = _oracle_attn
        assert X.install()
        from xfuser.core.distributed import (get_sequence_parallel_world_size, init_distributed_environment,
                                             initialize_model_parallel)
        init_distributed_environment(rank=rank, world_size=world)
        initialize_model_parallel(sequence_parallel_degree=world, ring_degree=1, ulysses_degree=world)  # flow/wanvideo.py:126
        assert get_sequence_parallel_world_size() == world
        cp = importlib.import_module("wan_ref.distributed.xdit_context_parallel")  # imports xfuser: the stand-in
        if variant == "dropin":
            sys.modules["videotuna.models.wan.wan.distributed.xdit_context_parallel"] = cp
            done = P.patch_sp()
            assert done["wan"] == 2 and getattr(cp.usp_dit_forward, "_b200vt_patched", False)
        m = _wan_model(model)
        for block in m.blocks:  # wan/text2video.py:266-270
            block.self_attn.forward = types.MethodType(cp.usp_attn_forward, block.self_attn)
        m.forward = types.MethodType(cp.usp_dit_forward, m)
        x, t, ctx, seq_len = _wan_inputs()
        with torch.no_grad():
            out = m(x, t, ctx, seq_len)
        q.put((rank, out[0].numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant", ["dropin", "reference_on_shim"])
def test_wan_usp_dit_forward_world2_equals_single_process_reference(variant):
    """variant "dropin": patch_sp() rebinds usp_dit_forward (and usp_attn_forward, which on CPU tensors raises Unsupported
    and hands over to the reference's) ; "reference_on_shim": the reference's wrappers unmodified over the xfuser stand-in."""
    _shims()
    model = _wan_modules()
    m = _wan_model(model)
    x, t, ctx, seq_len = _wan_inputs()
    with torch.no_grad():
        want = m(x, t, ctx, seq_len)[0]
    world = 2
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_wan_worker, args=(r, world, port, variant, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        torch.testing.assert_close(torch.from_numpy(res[r]), want, rtol=2e-4, atol=2e-5)


# ---------------------------------------------------------------------------------------------------------------------
# HunyuanVideo
# ---------------------------------------------------------------------------------------------------------------------
def _hy_modules():
    import importlib
    M = importlib.import_module("videotuna.models.hunyuan.hyvideo_i2v.modules.models")
    att = importlib.import_module("videotuna.models.hunyuan.hyvideo_i2v.modules.attenion")
    pos = importlib.import_module("videotuna.models.hunyuan.hyvideo_i2v.modules.posemb_layers")
    M.get_cu_seqlens = R.hunyuan_cu_seqlens  # the reference hard-codes device="cuda" (attenion.py:48)

    def attention_cpu(q_, k_, v_, cu_seqlens_q=None, **kw):  # mode="flash" semantics through the segment mask
        b, s = q_.shape[:2]
        mask = R.varlen_block_mask(cu_seqlens_q.tolist(), b * s)
        mask = torch.stack([mask[i * s:(i + 1) * s, i * s:(i + 1) * s] for i in range(b)])[:, None]
        return att.attention(q_, k_, v_, mode="torch", attn_mask=mask)

    M.attention = attention_cpu
    return M, pos


def _hy_model(M, cond):
    torch.manual_seed(3)
    m = M.HYVideoDiffusionTransformer(patch_size=[1, 2, 2], in_channels=4, hidden_size=128, heads_num=2,
                                      mlp_width_ratio=1.0, mm_double_blocks_depth=1, mm_single_blocks_depth=1,
                                      rope_dim_list=[8, 28, 28], text_projection="linear", text_states_dim=32,
                                      text_states_dim_2=16, i2v_condition_type=cond)
    with torch.no_grad():
        for p in m.parameters():
            if float(p.abs().max()) == 0.0:
                p.copy_(torch.randn_like(p) * 0.05)
    return m.eval()


def _hy_inputs(pos):
    g = torch.Generator().manual_seed(13)
    x = torch.randn(1, 4, 3, 8, 12, generator=g)               # 3 x 4 x 6 = 72 tokens; rows 4 and cols 6 divide by 2
    text = torch.randn(1, 10, 32, generator=g)
    mask = torch.zeros(1, 10, dtype=torch.long)
    mask[0, :7] = 1                                            # 3 padding tokens: the tail segment of cu_seqlens
    cos, sin = pos.get_nd_rotary_pos_embed([8, 28, 28], (3, 4, 6), theta=256, use_real=True, theta_rescale_factor=1)
    return x, torch.tensor([500.0]), text, mask, torch.randn(1, 16, generator=g), cos, sin


def _hy_worker(rank, world, port, cond, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        import b200vt.blocks as Bk
        import b200vt.patch as P
        import b200vt.sp as sp
        _shims()
        M, pos = _hy_modules()
        M.parallel_attention = Bk.hunyuan_parallel_attention  # what patch_blocks() installs (attenion.py:159-212)
        m = _hy_model(M, cond)
        pipe = types.SimpleNamespace(transformer=m)
        P.hunyuan_parallelize_transformer(pipe, attn_fn=_oracle_attn)  # CPU attention core for this test
        x, t, text, mask, text2, cos, sin = _hy_inputs(pos)
        with torch.no_grad():
            out = m(x, t, text, mask, text2, cos, sin, None, True)["x"]
        assert all(isinstance(b.hybrid_seq_parallel_attn, sp.UlyssesAttention) for b in list(m.double_blocks) + list(m.single_blocks))
        q.put((rank, out.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cond", [None, "token_replace"])
def test_hunyuan_parallelize_transformer_world2_equals_single_process_reference(cond):
    """The real hyvideo_i2v DiT (1 double + 1 single block) under b200vt.patch.hunyuan_parallelize_transformer on two ranks:
    latent split along the patch rows, RoPE tables sliced per rank, image shard + replicated valid text through Ulysses
    ("rear"), the padding tail attended locally, slabs gathered — equal to the unsharded reference forward."""
    _shims()
    M, pos = _hy_modules()
    if cond == "token_replace":
        pytest.skip("first-frame token count is taken from the LOCAL slab under SP (models.py:700-703): not a re-partitioning")
    m = _hy_model(M, cond)
    x, t, text, mask, text2, cos, sin = _hy_inputs(pos)
    with torch.no_grad():
        want = m(x, t, text, mask, text2, cos, sin, None, True)["x"]
    world = 2
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_hy_worker, args=(r, world, port, cond, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        torch.testing.assert_close(torch.from_numpy(res[r]), want, rtol=2e-4, atol=2e-5)


def test_install_ulysses_and_parallel_attention_padding_tail_single_process():
    """install_ulysses() sets the attribute on every block; hunyuan_parallel_attention (world size 1) equals the plain
    two-segment attention of the same tensors, padding tail included."""
    import types
    import b200vt.blocks as Bk
    import b200vt.patch as P
    import b200vt.sp as sp
    blocks = [types.SimpleNamespace(hybrid_seq_parallel_attn=None) for _ in range(3)]
    dit = types.SimpleNamespace(double_blocks=blocks[:1], single_blocks=blocks[1:])
    assert P.install_ulysses(dit) == 3 and all(isinstance(b.hybrid_seq_parallel_attn, sp.UlyssesAttention) for b in blocks)
    g = torch.Generator().manual_seed(2)
    L, T, valid, H, D = 24, 8, 5, 2, 16
    q, k, v = (torch.randn(1, L + T, H, D, generator=g, dtype=torch.float64) for _ in range(3))
    cu = torch.tensor([0, L + valid, L + T], dtype=torch.int32)
    out = Bk.hunyuan_parallel_attention(sp.UlyssesAttention(None, attn_fn=lambda a, b, c, s: R.sdpa_blhd(a, b, c, None, s)),
                                        q, k, v, img_q_len=L, img_kv_len=L, cu_seqlens_q=cu, cu_seqlens_kv=cu)
    want = R.hunyuan_attention_flash_semantics(q, k, v, cu)
    torch.testing.assert_close(out, want, rtol=1e-10, atol=1e-10)
