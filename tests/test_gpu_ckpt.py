"""Selective activation checkpointing (b200vt.ckpt): a checkpointed block keeps the attention outputs (O, LSE) instead of
re-running the attention forward in its backward pass. The output must equal that of full recomputation and of no
checkpointing bit for bit and the gradients to 1e-3 (same kernels on the same inputs; 128 tokens = one key tile, so that dQ has
no fp32 reduce-add whose order could differ between runs), and the attention forward must launch once, not twice."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Block(torch.nn.Module):
    def __init__(self, dim=256, heads=2):
        super().__init__()
        self.heads = heads
        self.qkv = torch.nn.Linear(dim, 3 * dim)
        self.proj = torch.nn.Linear(dim, dim)
        self.norm = torch.nn.LayerNorm(dim, elementwise_affine=False)

    def forward(self, x, shift, scale, gate):
        import b200vt.functional as Fn
        b, n, c = x.shape
        h = Fn.ln_modulate(x, None, None, scale, shift, 1e-6)
        q, k, v = self.qkv(h).view(b, n, 3, self.heads, c // self.heads).unbind(2)
        o = Fn.attention_blhd(q, k, v).reshape(b, n, c)
        return Fn.gate_residual(x, self.proj(o), gate)


def _close(a, b):
    # same kernels on the same inputs: equal in practice; the tolerance only absorbs a library GEMM choosing another split
    return float((a.float() - b.float()).abs().max()) <= 1e-3 * float(b.float().abs().max()) + 1e-7


def _run(mode):
    import torch.utils.checkpoint as tuc
    import b200vt._lib as L
    import b200vt.ckpt as CK
    torch.manual_seed(5)
    blk = _Block().cuda().bfloat16()
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(2, 128, 256, device="cuda", dtype=torch.bfloat16, generator=g).requires_grad_(True)
    vecs = [torch.randn(2, 256, device="cuda", dtype=torch.bfloat16, generator=g) * 0.1 for _ in range(3)]
    dy = torch.randn(2, 128, 256, device="cuda", dtype=torch.bfloat16, generator=g)
    L.profile_enable(True)
    y = x
    for _ in range(2):  # two checkpointed regions in a row
        if mode == "none":
            y = blk(y, *vecs)
        elif mode == "full":
            y = tuc.checkpoint(blk, y, *vecs, use_reentrant=False)
        elif mode == "selective":
            y = CK.checkpoint(blk, y, *vecs)
        else:  # the reference's call shape, picked up by the module-level wrapper
            y = torch.utils.checkpoint.checkpoint(blk, y, *vecs, use_reentrant=False)
    y.backward(dy)
    torch.cuda.synchronize()
    n_fwd, n_bwd = L.profile_read(L.K_ATTN_FWD)[1], L.profile_read(L.K_ATTN_BWD)[1]
    L.profile_enable(False)
    return [y.detach(), x.grad, *[p.grad for p in blk.parameters()]], n_fwd, n_bwd


def test_selective_checkpoint_keeps_attention_outputs():
    import torch.utils.checkpoint as tuc
    import b200vt.ckpt as CK
    ref, f0, b0 = _run("none")
    full, f1, b1 = _run("full")
    sel, f2, b2 = _run("selective")
    assert (f0, b0) == (2, 2) and (f1, b1) == (4, 2) and (f2, b2) == (2, 2)  # forward launches: recompute vs kept
    assert torch.equal(ref[0], full[0]) and torch.equal(ref[0], sel[0])  # forward output: bit for bit
    for a, b, c in zip(ref, full, sel):
        assert _close(a, b) and _close(a, c)
    CK.keep_attention_in_checkpoints()
    try:
        assert getattr(tuc.checkpoint, "_b200vt_wrapped", False)
        pat, f3, b3 = _run("patched")
        assert (f3, b3) == (2, 2)
        for a, b in zip(ref, pat):
            assert _close(a, b)
        # re-entrant calls and calls with their own context_fn pass through untouched
        blk = torch.nn.Linear(8, 8).cuda()
        z = tuc.checkpoint(blk, torch.randn(2, 8, device="cuda", requires_grad=True), use_reentrant=True)
        z.sum().backward()
    finally:
        CK.keep_attention_in_checkpoints(False)
    assert tuc.checkpoint is CK._ORIGINAL
