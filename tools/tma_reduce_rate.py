"""TMA fp32 reduce-add throughput (the backward kernel's dQ path): cycles per 16 KB box and aggregate TB/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vt._lib as L  # noqa: E402

iters = 2000
for n_tiles in (930,):  # 930 tiles x 64 KB = 61 MB: one head of K1's dQ accumulator (L2-resident)
    acc = torch.zeros(n_tiles * 128, 128, dtype=torch.float32, device="cuda")
    for blocks in (1, 148):
        for spread in (0, 1):
            for depth in (1, 2, 4):
                out = torch.zeros(blocks, dtype=torch.int64, device="cuda")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                L.call("vt_tma_reduce_rate", L.vp(acc.data_ptr()), n_tiles, iters, depth, spread, blocks,
                       L.vp(out.data_ptr()), None)
                e1.record()
                torch.cuda.synchronize()
                cyc = out.float().mean().item() / iters
                tbs = blocks * iters * 16384 / (e0.elapsed_time(e1) * 1e-3) / 1e12
                print(f"blocks={blocks:3d} spread={spread} depth={depth}: {cyc:7.1f} cycles per 16 KB reduce "
                      f"({16384 / cyc:5.1f} B/clk/SM), aggregate {tbs:5.2f} TB/s", flush=True)
