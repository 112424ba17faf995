"""Ulysses sequence parallelism on CPU: world_size-2 gloo processes against the single-process oracle.

SP correctness is defined as equality with the unsharded result: outputs and gradients of the sharded
run, gathered, must equal the oracle's attention on the full tensors. The CUDA kernel is not involved here — the
attention core is injected (the oracle's explicit softmax attention) so that only the host-side exchange logic
(pack/unpack, all-to-all, joint "rear" text tokens, autograd adjoints) is under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_ops as R


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_attn(q, k, v, softmax_scale):
    return R.sdpa_blhd(q, k, v, None, softmax_scale)


def _worker(rank, world, port, B, L, T, H, D, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import b200vt.sp as sp
        g = torch.Generator().manual_seed(1234)
        q, k, v = (torch.randn(B, L, H, D, generator=g, dtype=torch.float64) for _ in range(3))
        tq, tk, tv = (torch.randn(B, T, H, D, generator=g, dtype=torch.float64) for _ in range(3)) if T else (None,) * 3
        do = torch.randn(B, L // world + T, H, D, generator=torch.Generator().manual_seed(77 + rank), dtype=torch.float64)
        ql, kl, vl = (sp.shard_sequence(t).clone().requires_grad_(True) for t in (q, k, v))
        joint = {}
        if T:
            tql, tkl, tvl = (t.clone().requires_grad_(True) for t in (tq, tk, tv))
            joint = dict(joint_tensor_query=tql, joint_tensor_key=tkl, joint_tensor_value=tvl, joint_strategy="rear")
        attn = sp.UlyssesAttention(None, attn_fn=_oracle_attn)
        out = attn(None, ql, kl, vl, **joint)
        out.backward(do)
        res = {"out": out.detach(), "dq": ql.grad, "dk": kl.grad, "dv": vl.grad, "do": do}
        if T:
            res.update(dtq=tql.grad, dtk=tkl.grad, dtv=tvl.grad)
        result_q.put((rank, {n: t.numpy() for n, t in res.items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


# (1, 24, 6, 3, 8) and (1, 16, 0, 5, 8): head counts that do not divide the world size are padded with zero heads
@pytest.mark.parametrize("B,L,T,H,D", [(1, 24, 0, 4, 8), (1, 24, 6, 4, 8), (2, 16, 5, 2, 8), (1, 24, 6, 3, 8), (1, 16, 0, 5, 8)])
def test_ulysses_world2_matches_unsharded(B, L, T, H, D):
    world = 2
    ctx = mp.get_context("spawn")
    result_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, L, T, H, D, result_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(result_q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = {r: {n: torch.from_numpy(a) for n, a in d.items()} for r, d in results.items()}

    # unsharded oracle on the same tensors
    g = torch.Generator().manual_seed(1234)
    q, k, v = (torch.randn(B, L, H, D, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(3))
    if T:
        tq, tk, tv = (torch.randn(B, T, H, D, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(3))
        qa, ka, va = torch.cat([q, tq], 1), torch.cat([k, tk], 1), torch.cat([v, tv], 1)
    else:
        qa, ka, va = q, k, v
    ref = R.sdpa_blhd(qa, ka, va)
    Ls = L // world
    # the upstream gradient: image rows from each rank's shard; text rows are consumed on every rank -> sum
    do_full = torch.zeros_like(ref)
    for r in range(world):
        do_full[:, r * Ls:(r + 1) * Ls] = res[r]["do"][:, :Ls]
        if T:
            do_full[:, L:] += res[r]["do"][:, Ls:]
    ref.backward(do_full)

    for r in range(world):
        torch.testing.assert_close(res[r]["out"][:, :Ls], ref.detach()[:, r * Ls:(r + 1) * Ls], rtol=1e-10, atol=1e-10)
        if T:
            torch.testing.assert_close(res[r]["out"][:, Ls:], ref.detach()[:, L:], rtol=1e-10, atol=1e-10)
        for name, full in (("dq", q.grad), ("dk", k.grad), ("dv", v.grad)):
            torch.testing.assert_close(res[r][name], full[:, r * Ls:(r + 1) * Ls], rtol=1e-9, atol=1e-10)
    if T:
        # each rank holds the gradient of its own head slice of the replicated text tensors; their sum is the total
        for name, full in (("dtq", tq.grad), ("dtk", tk.grad), ("dtv", tv.grad)):
            total = sum(res[r][name] for r in range(world))
            torch.testing.assert_close(total, full, rtol=1e-9, atol=1e-10)


def test_single_process_degenerates_to_plain_attention():
    import b200vt.sp as sp
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn(1, 10, 2, 8, generator=g, dtype=torch.float64) for _ in range(3))
    tq, tk, tv = (torch.randn(1, 3, 2, 8, generator=g, dtype=torch.float64) for _ in range(3))
    out = sp.UlyssesAttention(None, attn_fn=_oracle_attn)(None, q, k, v, joint_tensor_query=tq, joint_tensor_key=tk,
                                                          joint_tensor_value=tv, joint_strategy="rear")
    ref = R.sdpa_blhd(torch.cat([q, tq], 1), torch.cat([k, tk], 1), torch.cat([v, tv], 1))
    torch.testing.assert_close(out, ref)
