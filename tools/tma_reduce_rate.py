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
            for depth in (2, 4, 18, 20):
                out = torch.zeros(blocks, dtype=torch.int64, device="cuda")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                L.call("vt_tma_reduce_rate", L.vp(acc.data_ptr()), n_tiles, iters, depth, spread, blocks,
                       L.vp(out.data_ptr()), None)
                e1.record()
                torch.cuda.synchronize()
                cyc = out.float().mean().item() / iters
                tbs = blocks * iters * 16384 / (e0.elapsed_time(e1) * 1e-3) / 1e12
                print(f"blocks={blocks:3d} spread={spread} {'1-D bulk' if depth >= 16 else 'tensor  '} depth={depth % 16}: {cyc:7.1f} cycles per 16 KB reduce "
                      f"({16384 / cyc:5.1f} B/clk/SM), aggregate {tbs:5.2f} TB/s", flush=True)

# reductions with concurrent TMA loads on the same SM
src = torch.zeros(930 * 128, 128, dtype=torch.bfloat16, device="cuda")
acc = torch.zeros(930 * 128, 128, dtype=torch.float32, device="cuda")
for blocks in (1, 148):
    for load_mode in (0, 1):
        out = torch.zeros(blocks, dtype=torch.int64, device="cuda")
        nld = torch.zeros(blocks, dtype=torch.int64, device="cuda")
        L.call("vt_tma_mixed_rate", L.vp(acc.data_ptr()), L.vp(src.data_ptr()), 930, iters, load_mode, blocks,
               L.vp(out.data_ptr()), L.vp(nld.data_ptr()), None)
        torch.cuda.synchronize()
        cyc = out.float().mean().item()
        print(f"mixed blocks={blocks:3d} loads={'on ' if load_mode else 'off'}: {cyc / iters:7.1f} cycles per 16 KB reduce; "
              f"{nld.float().mean().item() / iters:5.2f} 16 KB loads per reduce "
              f"({nld.float().mean().item() * 16384 / max(cyc, 1):5.1f} B/clk/SM loaded)", flush=True)
