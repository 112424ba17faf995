"""Whole-step CUDA graphs for the small-kernel denoisers.

The lvdm 3D-UNet (VideoCrafter / DynamiCrafter) issues ~11 000 kernel launches per LoRA finetune step at batch 2; a B200
finishes most of them faster than the host can issue the next, so in eager mode the step is bound by the launch path, not
by the device (device busy 241 ms of a 290 ms step). Every b200vt op launches on the current stream,
allocates its outputs and workspaces through the PyTorch caching allocator and never synchronises the host, so a complete
training step — forward, loss, backward, optimizer — can be captured ONCE into a CUDA graph and replayed: one launch per
step. This module is the small amount of ceremony that takes: warm-up on a side stream, capture, replay.

    step = b200vt.graph.GraphedStep(train_step)      # train_step(): reads static input tensors, returns a tensor (loss)
    for batch in loader:
        static_x.copy_(batch)                        # refill the static inputs in place
        loss = step()                                # one graph launch

Requirements on `fn` (those of torch.cuda.graph): static shapes and input tensors that are refilled in place; no host
synchronisation inside (no .item(), no Python branching on device values); a capturable optimizer
(torch.optim.AdamW(..., capturable=True)); activation checkpointing only with preserve_rng_state=False (saving the RNG
state is a host operation) — on 180 GB the lvdm UNet does not need it. Gradients must be None (zero_grad(set_to_none=True))
when the capture starts, so that the captured backward assigns them rather than accumulates.
Not used for the DiT backbones: their kernels are long and the eager step is already device-bound (tools/bench_denoiser.py
--graph: CogVideoX-2B 0.714 -> 0.709 s/it, Wan2.1-14B 1.286 -> 1.286).
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedStep:
    """Capture `fn()` into a CUDA graph after `warmup` eager runs on a side stream; calling the object replays it and
    returns fn's (static) result."""

    def __init__(self, fn: Callable[[], object], warmup: int = 3, pool=None):
        if not torch.cuda.is_available():
            raise RuntimeError("b200vt.graph.GraphedStep needs a CUDA device (there is no CPU path)")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # lazy initialisations (cuBLAS handles, kernel attributes, allocator pools) happen here
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.result = fn()

    def __call__(self):
        self.graph.replay()
        return self.result
