"""Decode the debug timeline of CTA (0,0,0) of the attention kernels (vt_debug_set_trace) and print per-iteration
deltas in cycles. Usage: python tools/trace_timeline.py bwd|fwd [L] [D]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the marks exist only in a -DVT_TRACE build: tools/build_variant.sh gpurun_tmp/lib_trace.so -DVT_TRACE
if "B200VT_LIB" not in os.environ and os.path.exists(os.path.join(ROOT, "gpurun_tmp", "lib_trace.so")):
    os.environ["B200VT_LIB"] = os.path.join(ROOT, "gpurun_tmp", "lib_trace.so")
import b200vt._lib as L  # noqa: E402
import b200vt.ops as ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "bwd"
Lq = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
H = 24
q, k, v, do = (torch.randn(1, Lq, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(4))
scale = 1 / math.sqrt(D)
o, lse = ops.attn_fwd(q, k, v, None, None, None, Lq, Lq, scale)
ops.attn_bwd(do, q, k, v, o, lse, None, None, None, Lq, Lq, scale)
torch.cuda.synchronize()
buf = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
L.call("vt_debug_set_trace", L.vp(buf.data_ptr()))
if which == "bwd":
    ops.attn_bwd(do, q, k, v, o, lse, None, None, None, Lq, Lq, scale)
else:
    ops.attn_fwd(q, k, v, None, None, None, Lq, Lq, scale)
torch.cuda.synchronize()
L.call("vt_debug_set_trace", None)
t = buf.cpu().view(4, 64, 8)
t0 = int(t[t > 0].min())
names = ({0: "compute", 1: "mma", 2: "drain", 3: "producer"} if which == "bwd" else
         {0: "softmax0", 1: "mma", 2: "softmax1", 3: "mma-fine"})
for it in range(20, 28):
    print(f"--- iteration {it}")
    for role in range(4):
        row = [int(x) - t0 if x > 0 else None for x in t[role, it]]
        if any(x is not None for x in row):
            print(f"  {names[role]:9s}", " ".join(f"{x:7d}" if x is not None else "      -" for x in row))
slot = 6 if which == "bwd" else 5
per = [int(t[0, i + 1, slot] - t[0, i, slot]) for i in range(16, 40) if t[0, i + 1, slot] > 0 and t[0, i, slot] > 0]
if per:
    print(f"period (role 0 slot {slot}):", per)
