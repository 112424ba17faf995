"""First-contact diagnostics on a B200: runs each building block in its own subprocess so a trap in one does not hide
the others, and prints a table. Usage: python tools/gpu_diag.py [probe|attn|all]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PROBE = r'''
import sys, torch
sys.path.insert(0, %r)
import b200vt.ops as ops, b200vt._lib as L
a_mode, b_mode, n = %d, %d, %d
g = torch.Generator().manual_seed(1)
a = torch.randn(128,128,generator=g).to(torch.bfloat16).cuda(); b = torch.randn(128,128,generator=g).to(torch.bfloat16).cuda()
kmaj=(16,1024,32); mn=(16384,1024,2048)
try:
    d = ops.umma_probe(a,b,a_mode,b_mode,n,a_desc=(mn if a_mode==1 else kmaj), b_desc=(mn if b_mode==1 else kmaj)); torch.cuda.synchronize()
except Exception as e:
    print("EXC", repr(e)[:200], "watchdog", L.watchdog()); sys.exit(1)
A = a.float() if a_mode != 1 else a.float().T
B = b.float()[:n].T if b_mode == 0 else b.float()[:, :n]
ref = A @ B
print("relerr %%.3e" %% float((d-ref).abs().max()/ref.abs().max()))
'''

ATTN = r'''
import sys, math, torch
sys.path.insert(0, %r)
import b200vt.ops as ops, b200vt._lib as L
B,Lq,Lk,H,D = %d,%d,%d,%d,%d
g = torch.Generator().manual_seed(2)
q = torch.randn(B,Lq,H,D,generator=g).to(torch.bfloat16); k = torch.randn(B,Lk,H,D,generator=g).to(torch.bfloat16); v = torch.randn(B,Lk,H,D,generator=g).to(torch.bfloat16)
try:
    o, lse = ops.attn_fwd(q.cuda(),k.cuda(),v.cuda(),None,None,None,Lq,Lk,1/math.sqrt(D)); torch.cuda.synchronize()
except Exception as e:
    print("EXC", repr(e)[:200], "watchdog", [hex(x) for x in L.watchdog()]); sys.exit(1)
s = torch.einsum("bihd,bjhd->bhij", q.float(), k.float())/math.sqrt(D)
ref = torch.einsum("bhij,bjhd->bihd", s.softmax(-1), v.float())
err = float((o.float().cpu()-ref).abs().max()/ref.abs().max())
lerr = float((lse.cpu()-torch.logsumexp(s,-1)).abs().max())
print("relerr %%.3e lse_abs_err %%.3e" %% (err, lerr))
'''


BWD = r'''
import sys, math, torch
sys.path.insert(0, %r)
import b200vt.functional as Fn, b200vt._lib as L
B,Lq,Lk,H,D = %d,%d,%d,%d,%d
g = torch.Generator().manual_seed(3)
mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16)
q, k, v, do = mk(B,Lq,H,D), mk(B,Lk,H,D), mk(B,Lk,H,D), mk(B,Lq,H,D)
qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q,k,v))
try:
    o = Fn.attention_blhd(qc,kc,vc); o.backward(do.cuda()); torch.cuda.synchronize()
except Exception as e:
    print("EXC", repr(e)[:200], "watchdog", [hex(x) for x in L.watchdog()]); sys.exit(1)
qr, kr, vr = (t.float().requires_grad_(True) for t in (q,k,v))
s = torch.einsum("bihd,bjhd->bhij", qr, kr)/math.sqrt(D)
ref = torch.einsum("bhij,bjhd->bihd", s.softmax(-1), vr); ref.backward(do.float())
def e(a,b): a=a.float().cpu(); return float((a-b).abs().max()/b.abs().max())
def c(a,b): a=a.float().cpu().flatten().double(); b=b.flatten().double(); return float(a@b/(a.norm()*b.norm()))
print("dq err %%.2e cos %%.5f | dk err %%.2e cos %%.5f | dv err %%.2e cos %%.5f" %% (e(qc.grad,qr.grad), c(qc.grad,qr.grad), e(kc.grad,kr.grad), c(kc.grad,kr.grad), e(vc.grad,vr.grad), c(vc.grad,vr.grad)))
'''


def run(code):
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
        if r.returncode != 0 and "EXC" not in out:
            out += " | rc=%d %s" % (r.returncode, r.stderr.strip()[-300:].replace("\n", " / "))
        return out
    except subprocess.TimeoutExpired:
        return "TIMEOUT"


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("probe", "all"):
        for a_mode, b_mode, n in [(0, 0, 128), (0, 0, 64), (0, 1, 128), (0, 1, 64), (2, 1, 128), (2, 0, 128), (1, 1, 128), (1, 0, 128)]:
            print(f"probe a_mode={a_mode} b_mode={b_mode} n={n}: {run(PROBE % (ROOT, a_mode, b_mode, n))}", flush=True)
    if what in ("attn", "all"):
        for cfg in [(1, 128, 128, 1, 128), (1, 128, 128, 1, 64), (1, 256, 256, 1, 128), (1, 256, 512, 2, 128), (2, 200, 333, 2, 128),
                    (1, 1000, 77, 3, 64), (1, 2560, 2560, 2, 64), (1, 4096, 4096, 2, 128)]:
            print(f"attn_fwd B,Lq,Lk,H,D={cfg}: {run(ATTN % ((ROOT,) + cfg))}", flush=True)
    if what in ("bwd", "all"):
        bwd()


def bwd():
    for cfg in [(1, 128, 128, 1, 128), (1, 128, 128, 1, 64), (1, 256, 256, 1, 128), (1, 256, 384, 2, 128), (2, 200, 333, 2, 128),
                (1, 1000, 77, 3, 64), (1, 2560, 2560, 2, 64), (1, 2048, 2048, 2, 128)]:
        print(f"attn_bwd B,Lq,Lk,H,D={cfg}: {run(BWD % ((ROOT,) + cfg))}", flush=True)


if __name__ == "__main__":
    main()
