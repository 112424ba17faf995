"""bench.py's multi-rank control flow on CPU (gloo, world size 2): every rank must enter every tools/bench_denoiser.py run —
each run's iterations contain collectives, so a rank that skips one (e.g. because only rank 0 holds the result records)
deadlocks the others. The denoiser tool is replaced by a stub whose run() performs a collective and returns a record on rank 0
only, exactly like the real one; a join timeout turns a deadlock into a failure."""
import os
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []
        stub = types.ModuleType("bench_denoiser")

        def run(ns, manage_dist=True, emit=True):
            calls.append((ns.model, bool(getattr(ns, "keep_attention", False))))
            t = torch.ones(1)
            dist.all_reduce(t)  # the real iteration's all-to-alls / all-reduces
            assert int(t.item()) == world
            if getattr(ns, "model", "") == "wan" and getattr(ns, "keep_attention", False):
                raise RuntimeError("synthetic failure of one variant, on every rank")
            if dist.get_rank() != 0:
                return None
            return {"s_per_it": 1.0, "steps": ns.steps, "warmup": ns.warmup, "peak_mem_GB": 1.0,
                    "config": {"parallelism": f"ulysses_sp{world}"}, "attention_share_of_step": 0.5,
                    "attention_tflops_in_step": 1000.0}

        stub.run = run
        sys.modules["bench_denoiser"] = stub
        sys.path.insert(0, ROOT)
        import bench
        out = bench.denoiser_it_s(world)
        dist.barrier()
        result[rank] = {"calls": calls, "keys": sorted(out)}
        if rank == 0:
            result["out"] = out
    finally:
        dist.destroy_process_group()


def test_every_rank_enters_every_denoiser_run():
    mgr = mp.Manager()
    result = mgr.dict()
    ctx = mp.spawn(_worker, args=(2, 29671, result), nprocs=2, join=False)
    import time
    deadline, done = time.time() + 180, False
    while not done and time.time() < deadline:  # join() returns once ONE more process has exited; True when all have
        done = ctx.join(timeout=5)
    if not done:
        for p in ctx.processes:
            p.kill()
    assert done, "bench.denoiser_it_s deadlocked: the ranks did not run the same sequence of iterations"
    want = [("hunyuan", False), ("hunyuan", True), ("wan", False), ("wan", True)]  # cogvideox: single-GPU only
    assert result[0]["calls"] == want and result[1]["calls"] == want
    assert result[1]["keys"] == []  # rank 0 alone holds the records
    out = result["out"]
    assert out["hunyuanvideo_720x1280x129f_lora"]["attention_outputs_kept"]["s_per_it_full_stack"] == 10.0
    assert "error" in out["wan2.1_t2v_14b_480x832x81f"]["attention_outputs_kept"]
    assert out["wan2.1_t2v_14b_480x832x81f"]["s_per_it_full_stack"] == 5.0


def _hung_process(q):
    sys.path.insert(0, ROOT)
    import json
    import time

    import bench
    bench.start_extras_watchdog(0, 1.0, lambda: print(json.dumps({"metric": "headline", "value": 1.0}), flush=True))
    q.put("armed")
    time.sleep(120)  # a hung collective, as far as Python can tell
    q.put("not reached")


def test_extras_watchdog_emits_the_headline_and_exits_zero(capfd):
    """bench.start_extras_watchdog: a process stuck after arming the watchdog prints the headline line and exits 0; a process
    that cancels it in time is left alone."""
    import time
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_hung_process, args=(q,))
    t0 = time.time()
    p.start()
    assert q.get(timeout=60) == "armed"
    p.join(timeout=30)
    assert p.exitcode == 0 and time.time() - t0 < 60
    assert '"metric": "headline"' in capfd.readouterr().out
    sys.path.insert(0, ROOT)
    import bench
    fired = []
    t = bench.start_extras_watchdog(1, 0.2, lambda: fired.append(1))  # rank != 0: 15 s later; cancelled long before
    t.cancel()
    time.sleep(0.5)
    assert not fired and not t.is_alive()
