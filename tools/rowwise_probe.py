"""Launch each memory-bound kernel a few times at its BASELINE shape (for `ncu --set full -k regex:...` captures).
    PROBE_ITERS=1 ncu --set full --clock-control none --import-source on -k regex:'ln_fwd|rmsnorm_rope_fwd|temporal_mma|groupnorm_fwd' \
        -o gpurun_out/rowwise python tools/rowwise_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vt.functional as Fn  # noqa: E402
import b200vt.ops as ops  # noqa: E402

N = int(os.environ.get("PROBE_ITERS", "3"))
x = torch.randn(1, 119056, 3072, device="cuda", dtype=torch.bfloat16)
sc = torch.randn(1, 3072, device="cuda") * 0.1
x.requires_grad_(True)
for _ in range(N):
    y = Fn.ln_modulate(x, sc, sc, eps=1e-6)
    y.backward(torch.randn_like(y))
    x.grad = None
xr = x.detach()
for _ in range(N):
    r = Fn.gate_residual(xr, y.detach(), sc)
del y, r
qkv = torch.randn(1, 118800, 3, 24, 128, device="cuda", dtype=torch.bfloat16)
w = torch.ones(128, device="cuda")
cs = torch.rand(118800, 128, device="cuda")
qv = qkv[:, :, 0].detach().requires_grad_(True)
for _ in range(N):
    z = Fn.qk_rmsnorm_rope(qv, w, cs, cs)
    z.backward(torch.randn_like(z))
    qv.grad = None
del z
q, k, v = (torch.randn(5120, 16, 5, 64, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
for _ in range(N):
    o = ops.temporal_attn_fwd(q, k, v, None, 0.125)
    o.backward(torch.randn_like(o))
gw = torch.ones(320, device="cuda")
for shape in ((32, 320, 40, 64), (2, 320, 16, 40, 64)):  # ResBlock / SpatialTransformer input; TemporalTransformer input
    xg = torch.randn(*shape, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    for _ in range(N):
        g = Fn.groupnorm_silu(xg, gw, gw, 32, 1e-5, silu=True)
        g.backward(torch.randn_like(g))
torch.cuda.synchronize()
print("ok")
