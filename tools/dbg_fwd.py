import sys, math, torch
sys.path.insert(0, '/root/repo')
import b200vt.ops as ops, b200vt._lib as L
for (B, Lq, Lk, H, D) in [(1,128,128,1,128),(1,256,256,1,128),(1,256,512,2,128),(1,128,128,1,64)]:
    q = torch.randn(B, Lq, H, D, device='cuda', dtype=torch.bfloat16)
    k = torch.randn(B, Lk, H, D, device='cuda', dtype=torch.bfloat16)
    v = torch.randn(B, Lk, H, D, device='cuda', dtype=torch.bfloat16)
    try:
        o, lse = ops.attn_fwd(q, k, v, None, None, None, Lq, Lk, 1/math.sqrt(D))
        torch.cuda.synchronize()
        ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1,2).float(), k.transpose(1,2).float(), v.transpose(1,2).float()).transpose(1,2)
        print((B,Lq,Lk,H,D), 'ok err', float((o.float()-ref).abs().max()/ref.abs().max()))
    except Exception as e:
        print((B,Lq,Lk,H,D), 'FAIL', str(e)[:100], 'watchdog', [hex(x) for x in L.watchdog()])
        break
