"""Launch the channels-last GroupNorm kernels a few times at the VideoCrafter2 level-0 shapes (for ncu captures):
    ncu --set full --clock-control none --import-source on -k regex:gn_nhwc -c 12 -o gpurun_out/gn_nhwc python tools/gn_nhwc_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vt.functional as Fn  # noqa: E402

gw = torch.ones(320, device="cuda")
for shape, fmt in (((32, 320, 40, 64), torch.channels_last), ((2, 320, 16, 40, 64), torch.channels_last_3d)):
    xs = [torch.randn(*shape, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=fmt).requires_grad_(True) for _ in range(6)]
    for x in xs:
        y = Fn.groupnorm_silu(x, gw, gw, 32, 1e-5, silu=True)
        y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("ok")
