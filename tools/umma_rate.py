"""tcgen05.mma issue-rate microbenchmark: cycles per 128 x N x 16 bf16 MMA by operand sourcing, with and without
concurrent shared-memory store traffic. Usage: python tools/umma_rate.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vt._lib as L  # noqa: E402

NAMES = {0: "SS K/K (S=QK^T)", 1: "TS A=TMEM, B MN (PV)", 2: "SS MN/MN (dQ)", 3: "SS K/MN (dK)"}
iters = 2000
for blocks in (148,):
    for mode in (0, 1, 2, 3):
        for n in (128, 64):
            for dep in (1, 0):
                for noise in (0, 4):
                    out = torch.zeros(blocks, dtype=torch.int64, device="cuda")
                    L.call("vt_umma_rate", mode | (0 if dep else 16), n, iters, noise, blocks, L.vp(out.data_ptr()), None)
                    torch.cuda.synchronize()
                    cyc = out.float().mean().item() / (iters * 8)
                    print(f"blocks={blocks:3d} mode={mode} {NAMES[mode]:22s} N={n:3d} accumulate-chain={dep} "
                          f"store-noise-warps={noise}: {cyc:6.1f} cycles/MMA", flush=True)
