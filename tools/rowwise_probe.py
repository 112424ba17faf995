import sys, torch
sys.path.insert(0, "/root/repo")
import b200vt.functional as Fn, b200vt.ops as ops, math
x = torch.randn(1, 119056, 3072, device="cuda", dtype=torch.bfloat16)
sc = torch.randn(1, 3072, device="cuda") * 0.1
for _ in range(3):
    y = Fn.ln_modulate(x, sc, sc, eps=1e-6)
qkv = torch.randn(1, 118800, 3, 24, 128, device="cuda", dtype=torch.bfloat16)
w = torch.ones(128, device="cuda"); cs = torch.rand(118800, 128, device="cuda")
for _ in range(3):
    z = Fn.qk_rmsnorm_rope(qkv[:, :, 0], w, cs, cs)
q = torch.randn(5120, 16, 5, 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    o = ops.temporal_attn_fwd(q, q, q, None, 0.125)
xg = torch.randn(32, 320, 40, 64, device="cuda", dtype=torch.bfloat16)
gw = torch.ones(320, device="cuda")
for _ in range(3):
    g = Fn.groupnorm_silu(xg, gw, gw, 32, 1e-5, silu=True)
torch.cuda.synchronize()
print("ok")
