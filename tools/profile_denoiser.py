"""Kernel-level breakdown of a DiT block-stack iteration (tools/bench_denoiser.py) with torch.profiler: device time per kernel
name. Usage: python tools/profile_denoiser.py --model hunyuan --double 1 --single 1 --optimizer adamw [bench_denoiser flags]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_denoiser as BD  # noqa: E402

args = BD.parse(sys.argv[1:])
args.steps, args.warmup = 1, 2
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    BD.run(args, manage_dist=False, emit=False)
    torch.cuda.synchronize()
from torch.autograd import DeviceType  # noqa: E402
agg = {}
for e in prof.events():
    if e.device_type == DeviceType.CUDA:
        t = agg.setdefault(e.name, [0.0, 0])
        t[0] += e.device_time_total / 1e3
        t[1] += 1
rows = sorted(((k, v[0], v[1]) for k, v in agg.items()), key=lambda r: -r[1])
total = sum(r[1] for r in rows)
print(f"model={args.model} device total {total:.1f} ms over 3 iterations (2 warm-up + 1) incl. setup kernels")
for k, ms, n in rows[:45]:
    print(f"{ms:10.2f} ms {100 * ms / total:5.1f} % {n:6d}  {k[:130]}")
