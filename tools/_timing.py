"""Device-time measurement shared by the tools: a kernel (or short op sequence) is timed WITHOUT the host's launch path.

`fn` is captured `inner` times into one CUDA graph (inner chosen so a replay lasts ~2 ms) and the graph replay is timed
with CUDA events; per-call time = replay time / inner. An eager `e0; fn(); e1` pair on an idle GPU would include the
20-40 us a torch.library op spends in the dispatcher before its kernel reaches the GPU — as long as or longer than the
small kernels themselves (GroupNorm, temporal attention: 30-80 us). If capture is refused (an op that synchronises), the
fallback queues the `inner` calls back to back, which hides the launch path only for kernels longer than it.
Callers that need cold inputs rotate buffers inside `fn`; the rotation is captured too."""
import torch


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def device_time_ms(fn, iters=10, target_ms=2.0, max_inner=64, use_graph=True):
    """Median per-call device milliseconds of fn()."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    fn()
    fn()
    e1.record()
    torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1) / 2, 1e-3)
    inner = int(max(2, min(max_inner, round(target_ms / est))))
    run, mode = None, "eager"
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(inner):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            run, mode = g.replay, "graph"
        except Exception:  # noqa: BLE001
            torch.cuda.synchronize()
            run = None
    if run is None:
        def run():
            for _ in range(inner):
                fn()
    ts = []
    for _ in range(iters):
        e0, e1 = _events()
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    device_time_ms.last_mode = mode
    return sorted(ts)[len(ts) // 2]


device_time_ms.last_mode = "eager"
