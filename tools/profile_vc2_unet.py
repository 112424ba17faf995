"""Kernel-level breakdown of one VideoCrafter2 UNet LoRA step (tools/bench_vc2_unet.py, `ours` or `torch` arm) with
torch.profiler: device time per kernel name, top 25. Usage: python tools/profile_vc2_unet.py [ours|torch] [--no-checkpoint]"""
import os
import sys
import types

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import bench_vc2_unet as VU  # noqa: E402

arm = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "ours"
args = types.SimpleNamespace(arm=arm, steps=1, warmup=3, no_checkpoint="--no-checkpoint" in sys.argv, check=False,
                             nchw="--nchw" in sys.argv, frozen_bf16="--frozen-bf16" in sys.argv)
VU.run(args, emit=False)  # builds, warms up
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes="--ops" in sys.argv) as prof:
    VU.run(types.SimpleNamespace(**{**vars(args), "warmup": 1, "steps": 2}), emit=False)
    torch.cuda.synchronize()
from torch.autograd import DeviceType  # noqa: E402
agg = {}
for e in prof.events():
    if e.device_type == DeviceType.CUDA:  # actual device activities (kernels, memcpys), not the host-side op ranges
        t = agg.setdefault(e.name, [0.0, 0])
        t[0] += e.device_time_total / 1e3 if hasattr(e, "device_time_total") else e.cuda_time_total / 1e3
        t[1] += 1
rows = sorted(((k, v[0], v[1]) for k, v in agg.items()), key=lambda r: -r[1])
total = sum(r[1] for r in rows)
layout = "nchw" if "--nchw" in sys.argv else "channels_last"
print(f"arm={arm} layout={layout if arm == 'ours' else 'nchw'} checkpoint={not args.no_checkpoint}  device total {total:.1f} ms over 2 steps")
groups = {"conv / gemm (cudnn, cutlass, cublas)": ("cutlass", "cudnn", "gemm", "sm90", "sm100", "nvjet", "xmma"),
          "layout conversion (nchw<->nhwc)": ("nchwToNhwc", "nhwcToNchw"),
          "b200vt kernels": ("vt::",),
          "torch elementwise / copies": ("at::native",)}
for gname, pats in groups.items():
    ms = sum(r[1] for r in rows if any(p_ in r[0] for p_ in pats) and not (gname.startswith("conv") and ("nchwToNhwc" in r[0] or "nhwcToNchw" in r[0])))
    print(f"  {ms:9.2f} ms  {gname}")
for k, ms, n in rows[:60]:
    print(f"{ms:10.2f} ms {n:6d}  {k[:120]}")
if "--ops" in sys.argv:  # which torch ops the remaining stock kernels come from (self device time, grouped by input shapes)
    print("\n== torch ops by self device time (top 70, with input shapes)")
    ka = prof.key_averages(group_by_input_shape=True)
    ops = sorted((e for e in ka if e.key.startswith(("aten::", "b200vt")) or "Backward" in e.key), key=lambda e: -e.self_device_time_total)[:70]
    for e in ops:
        print(f"{e.self_device_time_total / 1e3:10.2f} ms {e.count:6d}  {e.key[:40]:40s} {str(e.input_shapes)[:110]}")
