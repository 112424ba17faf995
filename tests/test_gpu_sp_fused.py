"""Ulysses attention with the output exchange fused into the attention kernel's epilogue (peer stores over NVLink through
torch symmetric memory): two ranks on two GPUs must reproduce the single-GPU result — outputs and every gradient — and
agree with the NCCL all-to-all path. Needs >= 2 GPUs; skipped otherwise (the 1-GPU runs cover everything else)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, fused, T, result):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), B200VT_SP_FUSED="1" if fused else "0")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import b200vt.functional as Fn
        import b200vt.sp as sp
        L, H, D = 1024, 4, 128
        g = torch.Generator(device="cuda").manual_seed(1234)
        full = [torch.randn(1, L, H, D, device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(3)]
        txt = [torch.randn(1, T, H, D, device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(3)] if T else None
        d_out = torch.randn(1, L + T, H, D, device="cuda", dtype=torch.bfloat16, generator=g)
        # single-GPU reference on the full tensors (same kernels, no exchange)
        ref_in = [(torch.cat([f, t], 1) if T else f).clone().requires_grad_(True) for f, t in zip(full, txt or full)]
        ref = Fn.attention_blhd(*ref_in)
        ref.backward(d_out)
        S = L // world
        sl = slice(rank * S, (rank + 1) * S)
        loc = [f[:, sl].clone().requires_grad_(True) for f in full]
        tloc = [t.clone().requires_grad_(True) for t in txt] if T else [None] * 3
        attn = sp.UlyssesAttention()
        kw = dict(joint_tensor_query=tloc[0], joint_tensor_key=tloc[1], joint_tensor_value=tloc[2], joint_strategy="rear") if T else {}
        for _ in range(2):  # twice: the symmetric buffers are reused across calls
            for t in loc + [x for x in tloc if x is not None]:
                t.grad = None
            out = attn(None, *loc, **kw)
            # the text rows' output is replicated on every rank: each replica carries 1/world of its gradient, so the
            # sum over ranks equals the single-GPU gradient
            d_loc = torch.cat([d_out[:, sl], d_out[:, L:] / world], 1) if T else d_out[:, sl]
            out.backward(d_loc)
        want = torch.cat([ref[:, sl], ref[:, L:]], 1) if T else ref[:, sl]
        errs = {"out": float((out.float() - want.float()).abs().max())}
        for n, a, b in zip("qkv", loc, ref_in):
            errs["d" + n] = float((a.grad.float() - b.grad[:, sl].float()).abs().max() / b.grad.float().abs().max())
        if T:
            # Replicated text rows: a rank's copy receives the gradient of its own head slice only (zeros elsewhere), and
            # every rank applied 1/world of the text rows' upstream gradient, which the head gather's adjoint sums at the
            # owner — so the rank-summed gradient equals the single-GPU gradient of the text rows (as test_sp_gloo.py checks
            # on CPU).
            for n, a, b in zip("qkv", tloc, ref_in):
                gsum = a.grad.float().clone()
                dist.all_reduce(gsum)
                want_g = b.grad[:, L:].float()
                errs["dt" + n] = float((gsum - want_g).abs().max() / want_g.abs().max())
        result[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("T", [0, 64])
def test_ulysses_two_gpus_matches_single_gpu(fused, T):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = 29600 + (2 if fused else 0) + (1 if T else 0)
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_worker, args=(2, port, fused, T, result), nprocs=2, join=True)
    for rank in (0, 1):
        e = result[rank]
        assert e["out"] <= 2e-2, (rank, e)
        assert e["dq"] <= 2e-2 and e["dk"] <= 2e-2 and e["dv"] <= 2e-2, (rank, e)
        if T:
            assert e["dtq"] <= 2e-2 and e["dtk"] <= 2e-2 and e["dtv"] <= 2e-2, (rank, e)


def _wan_worker(rank, world, port, result):
    """Wan self-attention, sequence parallel (patch.wan_usp_attn_forward bound like wan/text2video.py:261-271 binds
    usp_attn_forward) against the unsharded drop-in on the same weights: output slice, input gradient slice and the
    rank-summed weight gradients."""
    import types

    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import b200vt.patch as P
        from helpers import WanAttnShell
        from oracle import ref_ops as R
        dim, heads, grid = 1024, 8, (4, 8, 16)  # head dim 128; 512 tokens, 256 per rank
        L = grid[0] * grid[1] * grid[2]
        torch.manual_seed(7)
        attn = WanAttnShell(dim, heads, cross=False)
        with torch.no_grad():
            for n_, p_ in attn.named_parameters():
                p_.copy_(torch.randn_like(p_) * (0.03 if p_.dim() == 2 else 0.1) + (1.0 if "norm" in n_ else 0.0))
        attn = attn.to("cuda", torch.bfloat16)
        g = torch.Generator(device="cuda").manual_seed(99)
        x = torch.randn(1, L, dim, device="cuda", dtype=torch.bfloat16, generator=g)
        d_out = torch.randn(1, L, dim, device="cuda", dtype=torch.bfloat16, generator=g)
        freqs = R.wan_freqs_table(dim // heads).cuda()
        grid_sizes = torch.tensor([list(grid)])
        seq_lens = torch.tensor([L], device="cuda")
        xr = x.clone().requires_grad_(True)
        ref = attn(xr, seq_lens, grid_sizes, freqs)  # single-GPU drop-in (blocks.wan_self_attention_forward)
        ref.backward(d_out)
        ref_wgrad = {n_: p_.grad.float().clone() for n_, p_ in attn.named_parameters()}
        for p_ in attn.parameters():
            p_.grad = None
        S = L // world
        sl = slice(rank * S, (rank + 1) * S)
        xs = x[:, sl].clone().requires_grad_(True)
        sp_forward = types.MethodType(P.wan_usp_attn_forward, attn)
        out = sp_forward(xs, seq_lens, grid_sizes, freqs)
        out.backward(d_out[:, sl])
        errs = {"out": float((out.float() - ref[:, sl].float()).abs().max() / ref.float().abs().max()),
                "dx": float((xs.grad.float() - xr.grad[:, sl].float()).abs().max() / xr.grad.float().abs().max())}
        worst = 0.0
        for n_, p_ in attn.named_parameters():
            gsum = p_.grad.float().clone()
            dist.all_reduce(gsum)
            worst = max(worst, float((gsum - ref_wgrad[n_]).abs().max() / ref_wgrad[n_].abs().max().clamp_min(1e-20)))
        errs["dw"] = worst
        result[rank] = errs
    finally:
        dist.destroy_process_group()


def test_wan_usp_attention_two_gpus_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_wan_worker, args=(2, 29611, result), nprocs=2, join=True)
    for rank in (0, 1):
        e = result[rank]
        assert e["out"] <= 2e-2 and e["dx"] <= 2e-2 and e["dw"] <= 3e-2, (rank, e)


def _ckpt_worker(rank, world, port, result):
    """Selective checkpointing (b200vt.ckpt) around sequence-parallel attention with the fused exchange: the op that holds the
    symmetric-memory barriers, the scatter kernel and the copy-out is kept as a whole, so a checkpointed region's backward
    re-runs neither — on every rank alike — and outputs (bit for bit) / gradients (to rounding) equal the un-checkpointed run."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), B200VT_SP_FUSED="1")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import b200vt._lib as L
        import b200vt.ckpt as CK
        import b200vt.sp as sp
        S, H, D, T = 512, 4, 128, 64
        g = torch.Generator(device="cuda").manual_seed(100 + rank)
        loc = [torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(3)]
        g0 = torch.Generator(device="cuda").manual_seed(7)
        txt = [torch.randn(1, T, H, D, device="cuda", dtype=torch.bfloat16, generator=g0) for _ in range(3)]
        d_out = torch.randn(1, S + T, H, D, device="cuda", dtype=torch.bfloat16, generator=g)
        lin = torch.nn.Linear(D, D).cuda().bfloat16()
        attn = sp.UlyssesAttention()

        def region(q, k, v, tq, tk, tv):
            o = attn(None, lin(q), k, v, joint_tensor_query=tq, joint_tensor_key=tk, joint_tensor_value=tv, joint_strategy="rear")
            return lin(o)

        outs = {}
        for mode in ("none", "selective"):
            ins = [t.clone().requires_grad_(True) for t in loc + txt]
            lin.zero_grad(set_to_none=True)
            L.profile_enable(True)
            y = region(*ins) if mode == "none" else CK.checkpoint(region, *ins)
            y.backward(d_out)
            torch.cuda.synchronize()
            n_fwd = L.profile_read(L.K_ATTN_FWD)[1]
            L.profile_enable(False)
            outs[mode] = ([y.detach()] + [t.grad for t in ins] + [lin.weight.grad.clone()], n_fwd)
        # (dQ is accumulated over the key tiles by fp32 reduce-adds in arrival order: equal to rounding, not bit for bit)
        same = torch.equal(outs["none"][0][0], outs["selective"][0][0]) and all(
            float((a.float() - b.float()).abs().max()) <= 1e-2 * float(b.float().abs().max()) + 1e-6
            for a, b in zip(outs["none"][0], outs["selective"][0]))
        result[rank] = {"same": bool(same), "fwd_launches": (outs["none"][1], outs["selective"][1])}
    finally:
        dist.destroy_process_group()


def test_selective_checkpoint_with_fused_exchange_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_ckpt_worker, args=(2, 29621, result), nprocs=2, join=True)
    for rank in (0, 1):
        assert result[rank]["same"], result[rank]
        assert result[rank]["fwd_launches"] == (1, 1), result[rank]  # kept: no second forward launch in the backward
