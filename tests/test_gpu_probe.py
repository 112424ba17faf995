"""tcgen05/TMA building blocks: one 128 x n x 128 tile with every operand source the attention kernels use."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(128, 128, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(128, 128, generator=g).to(torch.bfloat16).cuda()
    return a, b


def _expected(a, b, a_mode, b_mode, n):
    A = a.float() if a_mode != 1 else a.float().T          # a_mode 1: buffer holds A^T (K x M)
    Bm = b.float()[:n].T if b_mode == 0 else b.float()[:, :n]  # b_mode 0: (N x K); 1: (K x N)
    return A @ Bm


@pytest.mark.parametrize("a_mode,b_mode,n", [(0, 0, 128), (0, 0, 64), (0, 1, 128), (0, 1, 64), (2, 1, 128), (2, 1, 64),
                                             (2, 0, 128), (1, 1, 128), (1, 0, 128)])
def test_umma_probe(a_mode, b_mode, n):
    import b200vt._lib as L
    import b200vt.ops as ops
    if not L.has_experiments():
        pytest.skip("vt_umma_probe exists only in -DVT_EXPERIMENTS builds (tools/build_variant.sh + B200VT_LIB)")
    a, b = _mk(a_mode * 10 + b_mode)
    kmaj = (16, 1024, 32)          # K-major SW128: LBO ignored, SBO = 8 rows * 128 B, 32 B per 16-element k-step
    mnmaj = (16384, 1024, 2048)    # MN-major SW128: LBO = next 64-wide box, SBO = next 8 k-rows, 16 k-rows per step
    d = ops.umma_probe(a, b, a_mode, b_mode, n, a_desc=(mnmaj if a_mode == 1 else kmaj),
                       b_desc=(mnmaj if b_mode == 1 else kmaj))
    torch.cuda.synchronize()
    ref = _expected(a, b, a_mode, b_mode, n)
    err = float((d - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, f"a_mode={a_mode} b_mode={b_mode} n={n}: rel err {err}"
