#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark: HunyuanVideo 720x1280x129-frame joint text+video attention,
forward + backward, bf16 (BASELINE.json metric "3D-attn fwd+bwd TFLOP/s"; workload K1 of BASELINE.md §3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one forward + one backward of the attention call one HunyuanVideo block makes
(reference: attention(q, k, v, mode="flash", cu_seqlens...) at hyvideo_t2v/modules/models.py:204, under N > 1 GPUs
parallel_attention(hybrid_seq_parallel_attn, ...) at :215) on synthetic q, k, v, dO of shape (1, 118800 + 256, 24, 128).
    value   whole-job TFLOP/s with inputs resident in HBM (14 * B * H * L^2 * D algorithmic FLOPs per step)
    e2e     the same metric through the public Python API with HOST (pinned) buffers: host->device copies of q, k, v, dO
            and device->host copies of out, dq, dk, dv inside the timed region
    roofline  the dominant kernel (5-GEMM backward) timed live with CUDA events on its launch stream (vt_profile_*)
    cpu_baseline  the oracle's restatement of the reference's mode="torch" path (F.scaled_dot_product_attention) on
            the host cores, on a bounded sample of the same workload
N > 1 (launched by torchrun, one rank per GPU): Ulysses sequence parallelism — image tokens sharded over ranks, text
tokens replicated ("rear" joint strategy), NCCL all-to-all before and after the local kernel; strong scaling.
--impl reference: the reference's own CPU attention path timed on the host cores (rank 0 only).
Further keys on the same line (each measured in this process, with its own clocks record; --no-extras skips them):
    sp_parity          N > 1: one sequence-parallel step compared, for 2 heads, with the unsharded kernels run on the
                       gathered tensors (outputs and all gradients; must be <= 2e-2)
    library_baselines  N = 1: cuDNN SDPA (what the reference's mode="torch" reaches on torch 2.11) and flash-attn 2
                       (mode="flash") forward / backward on K1 and K4, next to the b200vt kernels, same inputs
    denoiser_it_s      the metric's second half: forward + backward of DiT block stacks through tools/bench_denoiser.py
                       (HunyuanVideo 2 double + 4 single blocks = 1/10 of the stack, Wan2.1-14B 8 of 40 blocks, both
                       Ulysses at N GPUs; CogVideoX-2B all 30 blocks at N = 1)
    rowwise            N = 1: achieved GB/s and fraction of the measured copy bandwidth of the memory-bound kernels
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ---- workload K1 -----------------------------------------------------------------
IMG_TOKENS = 33 * 45 * 80      # 129 frames -> 33 latent frames; 720x1280 -> 45 x 80 patches
TXT_TOKENS = 256
HEADS, HEAD_DIM = 24, 128
SEQ = IMG_TOKENS + TXT_TOKENS  # 119 056
SEED = 20230211                # the reference's default --seed (scripts/train_new.py:34)
WORKLOAD = "hunyuanvideo_720x1280x129f_attention_fwd_bwd"
METRIC = "attn_fwd_bwd_tflops"
UNIT = "TFLOP/s"


def flops_fwd(lq: int, lk: int, h: int = HEADS, d: int = HEAD_DIM, b: int = 1) -> float:
    return 4.0 * b * h * lq * lk * d


def flops_fwd_bwd(lq: int, lk: int, h: int = HEADS, d: int = HEAD_DIM, b: int = 1) -> float:
    return 3.5 * flops_fwd(lq, lk, h, d, b)  # bwd = 2.5 x fwd; recompute is not counted


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# =====================================================================================================================
# clocks
# =====================================================================================================================
class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="b200vt_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001  (no nvidia-smi: report nothing rather than fail the bench)
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for line in fh:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0]))
                    smax.append(float(f[1]))
                    power.append(float(f[2]))
                except ValueError:
                    continue
                for n, val in zip(names, f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# =====================================================================================================================
# CPU legs (oracle port of the reference's torch path) — the only place bench.py executes oracle/
# =====================================================================================================================
REFERENCE_TREE = "/root/reference"
_CPU_ATTENTION = {}


def cpu_attention_fn():
    """(callable, kind): the reference's own `attention(q, k, v, mode="torch")` (hyvideo_t2v/modules/attenion.py:60-156)
    imported from the reference tree when it is present (development container; kind "reference"), else the oracle's
    restatement of the same call (the GPU box has no /root/reference; kind "port")."""
    if not _CPU_ATTENTION:
        fn, kind = None, "port"
        if os.path.isdir(os.path.join(REFERENCE_TREE, "videotuna")):
            try:
                sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
                import importlib
                import make_golden
                make_golden.install_shims()  # import shims for absent wheels (no arithmetic)
                att = importlib.import_module("videotuna.models.hunyuan.hyvideo_t2v.modules.attenion")
                fn, kind = (lambda q, k, v: att.attention(q, k, v, mode="torch")), "reference"
            except Exception:  # noqa: BLE001
                fn = None
        if fn is None:
            from oracle import ref_ops as R
            fn, kind = R.hunyuan_attention_torch_fused, "port"
        _CPU_ATTENTION.update(fn=fn, kind=kind)
    return _CPU_ATTENTION["fn"], _CPU_ATTENTION["kind"]


def _cpu_sample_step(q, k, v, do):
    """One fwd+bwd of the reference's mode="torch" attention on host tensors (B, Lq|Lk, H, D)."""
    import torch
    out = cpu_attention_fn()[0](q, k, v)
    torch.autograd.grad(out, (q, k, v), do)


def _cpu_make(lq: int, lk: int, heads: int, dtype):
    import torch
    g = torch.Generator().manual_seed(SEED)
    q = torch.randn(1, lq, heads, HEAD_DIM, generator=g).to(dtype).requires_grad_(True)
    k = torch.randn(1, lk, heads, HEAD_DIM, generator=g).to(dtype).requires_grad_(True)
    v = torch.randn(1, lk, heads, HEAD_DIM, generator=g).to(dtype).requires_grad_(True)
    do = torch.randn(1, lq, heads * HEAD_DIM, generator=g).to(dtype)
    return q, k, v, do


def cpu_size_sample(target_s: float):
    """Pick a bounded sample of K1 — 1 of 24 heads (heads are independent), all 119 056 keys, a prefix of the query
    rows — whose fwd+bwd takes about `target_s` on this host. Returns (lq, lk, heads, est TFLOP/s)."""
    import torch
    torch.set_num_threads(host_threads())
    lq0 = 2048
    q, k, v, do = _cpu_make(lq0, 16384, 1, torch.bfloat16)
    _cpu_sample_step(q, k, v, do)
    t0 = time.perf_counter()
    _cpu_sample_step(q, k, v, do)
    dt = time.perf_counter() - t0
    rate = flops_fwd_bwd(lq0, 16384, 1) / dt  # FLOP/s
    lq = int(target_s * rate / (14.0 * SEQ * HEAD_DIM))
    lq = max(1024, min(SEQ, (lq // 1024) * 1024))
    return lq, SEQ, 1, rate / 1e12


def cpu_time_sample(lq: int, lk: int, heads: int, steps: int, warmup: int):
    import torch
    torch.set_num_threads(host_threads())
    q, k, v, do = _cpu_make(lq, lk, heads, torch.bfloat16)
    for _ in range(warmup):
        _cpu_sample_step(q, k, v, do)
    t0 = time.perf_counter()
    for _ in range(steps):
        _cpu_sample_step(q, k, v, do)
    dt = (time.perf_counter() - t0) / steps
    return flops_fwd_bwd(lq, lk, heads) / dt / 1e12, dt


def sample_text(lq, lk, heads):
    return (f"{heads} of {HEADS} heads (heads are independent), first {lq} of {SEQ} query rows against all {lk} keys, "
            f"bf16, fwd+bwd through F.scaled_dot_product_attention + autograd")


def run_reference(args):
    """--impl reference: the reference's own CPU attention path (hunyuan attention(mode='torch') =
    F.scaled_dot_product_attention, attenion.py:101-106) on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    lq, lk, heads, _ = cpu_size_sample(target_s=6.0)
    tflops, dt = cpu_time_sample(lq, lk, heads, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(tflops, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B": 1, "L": SEQ, "H": HEADS, "D": HEAD_DIM,
                   "note": "each step is a bounded sample of the workload; TFLOP/s is size-independent"},
        "cpu_baseline": {"value": round(tflops, 4), "unit": UNIT, "cores": host_threads(), "kind": cpu_attention_fn()[1],
                         "sample": sample_text(lq, lk, heads)},
        "e2e": {"value": round(tflops, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# GPU legs
# =====================================================================================================================
OTHER_SHAPES = {  # (B, Lq, Lk, H, D)
    "k2_wan14b_480x832x81_self": (1, 32760, 32760, 40, 128),
    "k3_cogvideox2b_480x720x49": (1, 17776, 17776, 30, 64),
    "k4_videocrafter2_spatial_self_level0_b2": (32, 2560, 2560, 5, 64),
}


def other_configs(torch, dev, iters: int = 5):
    """Attention fwd+bwd TFLOP/s of the remaining configurations through the same ops (CUDA events, median of `iters`)."""
    import math

    import b200vt.ops as ops
    out = {}
    for name, (B, Lq, Lk, H, D) in OTHER_SHAPES.items():
        g = torch.Generator(device=dev).manual_seed(SEED)
        q, k, v, do = (torch.randn(B, n, H, D, device=dev, dtype=torch.bfloat16, generator=g) for n in (Lq, Lk, Lk, Lq))
        scale = 1.0 / math.sqrt(D)

        def step():
            o, lse = ops.attn_fwd(q, k, v, None, None, None, Lq, Lk, scale)
            ops.attn_bwd(do, q, k, v, o, lse, None, None, None, Lq, Lk, scale)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        # short steps (K4: ~1.2 ms) are queued back to back so that the host's launch path (~50 us per op) stays off the
        # device timeline: ~20 ms of work per event pair
        inner = max(1, min(20, int(round(20.0 / max(e0.elapsed_time(e1), 1e-3)))))
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(inner):
                step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
        ms = sorted(ts)[len(ts) // 2]
        out[name] = {"ms_fwd_bwd": round(ms, 3), "tflops": round(flops_fwd_bwd(Lq, Lk, H, D, B) / (ms * 1e-3) / 1e12, 1),
                     "steps_per_event_pair": inner}
        del q, k, v, do
        torch.cuda.empty_cache()
    return out


def _ev_time(torch, fn, iters: int, warmup: int = 1) -> float:
    """Device ms per call of fn: CUDA events around `iters` back-to-back calls after `warmup` untimed ones."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def library_baselines(torch, dev):
    """cuDNN SDPA (F.scaled_dot_product_attention with the cuDNN backend: what the reference's mode="torch" / diffusers
    processors reach on torch 2.11) and flash-attn 2 (the reference's mode="flash", attenion.py:108-123) forward and
    backward on K1 and K4, next to the b200vt kernels: same process, same seeded inputs, device time from CUDA events."""
    import math

    import torch.nn.functional as F

    import b200vt.ops as ops
    shapes = {"k1_hunyuan": (1, SEQ, SEQ, HEADS, HEAD_DIM, 2), "k4_videocrafter2_spatial_self_level0_b2": (32, 2560, 2560, 5, 64, 10)}
    out = {}
    for name, (B, Lq, Lk, H, D, iters) in shapes.items():
        g = torch.Generator(device=dev).manual_seed(SEED)
        q, k, v, do = (torch.randn(B, n, H, D, device=dev, dtype=torch.bfloat16, generator=g) for n in (Lq, Lk, Lk, Lq))
        scale = 1.0 / math.sqrt(D)
        ff, fb = flops_fwd(Lq, Lk, H, D, B), 2.5 * flops_fwd(Lq, Lk, H, D, B)
        rec = {}

        def tf(flops, ms):
            return round(flops / (ms * 1e-3) / 1e12, 1)

        # ---- b200vt ----
        o, lse = ops.attn_fwd(q, k, v, None, None, None, Lq, Lk, scale)
        f_ms = _ev_time(torch, lambda: ops.attn_fwd(q, k, v, None, None, None, Lq, Lk, scale), iters)
        b_ms = _ev_time(torch, lambda: ops.attn_bwd(do, q, k, v, o, lse, None, None, None, Lq, Lk, scale), iters)
        rec["b200vt"] = {"fwd_ms": round(f_ms, 3), "bwd_ms": round(b_ms, 3), "fwd_tflops": tf(ff, f_ms), "bwd_tflops": tf(fb, b_ms),
                         "fwd_bwd_tflops": tf(ff + fb, f_ms + b_ms)}
        del o, lse
        # ---- library kernels through autograd (forward timed alone; backward = retained-graph grad) ----
        ql, kl, vl = (t.detach().requires_grad_(True) for t in (q, k, v))

        def lib_arm(label, fwd):
            try:
                y = fwd()
                f_ms = _ev_time(torch, fwd, iters)
                b_ms = _ev_time(torch, lambda: torch.autograd.grad(y, (ql, kl, vl), do, retain_graph=True), iters)
                rec[label] = {"fwd_ms": round(f_ms, 3), "bwd_ms": round(b_ms, 3), "fwd_tflops": tf(ff, f_ms),
                              "bwd_tflops": tf(fb, b_ms), "fwd_bwd_tflops": tf(ff + fb, f_ms + b_ms)}
                del y
            except Exception as e:  # noqa: BLE001  (a library that refuses the shape is a result, not a bench failure)
                rec[label] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
            torch.cuda.empty_cache()

        def cudnn_fwd():
            from torch.nn.attention import SDPBackend, sdpa_kernel
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                return F.scaled_dot_product_attention(ql.transpose(1, 2), kl.transpose(1, 2), vl.transpose(1, 2)).transpose(1, 2)

        def fa2_fwd():
            from flash_attn import flash_attn_func
            return flash_attn_func(ql, kl, vl)

        lib_arm("cudnn_sdpa", cudnn_fwd)
        lib_arm("flash_attn_2", fa2_fwd)
        rec["shape_BLqLkHD"] = [B, Lq, Lk, H, D]
        out[name] = rec
        del q, k, v, do, ql, kl, vl
        torch.cuda.empty_cache()
    try:
        import flash_attn
        out["versions"] = {"flash_attn": flash_attn.__version__, "cudnn": torch.backends.cudnn.version(), "torch": torch.__version__}
    except Exception:  # noqa: BLE001
        pass
    return out


def start_extras_watchdog(rank: int, limit_s: float, emit_headline):
    """A daemon timer for the multi-rank extras: if they have not finished (and cancelled it) after `limit_s`, rank 0 calls
    `emit_headline()` — which prints the json line from the values measured before the extras — and every rank leaves with
    exit code 0 (the others 15 s later, so that rank 0's line is out first). A hung collective cannot be interrupted from
    Python; leaving the process is the only way to hand the driver a line instead of a timeout."""
    import threading

    def fire():
        if rank == 0:
            try:
                emit_headline()
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"error": f"watchdog: {type(e).__name__}: {str(e)[:200]}"}), flush=True)
        sys.stdout.flush()
        os._exit(0)

    t = threading.Timer(limit_s + (0.0 if rank == 0 else 15.0), fire)
    t.daemon = True
    t.start()
    return t


def denoiser_it_s(world: int):
    """Finetuning iterations (forward + backward + AdamW step) of DiT block stacks (tools/bench_denoiser.py code path,
    `ours` arm: the reference blocks' constructors with the b200vt drop-in forwards, per-block activation checkpointing,
    3 warm-ups, 2 timed iterations)."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_denoiser as BD
    runs = [("hunyuanvideo_720x1280x129f_lora", dict(model="hunyuan", double=2, single=4, layers=None), 10.0,
             "2 of 20 double + 4 of 40 single blocks = exactly 1/10 of the block stack at the full 119 056 tokens"),
            ("wan2.1_t2v_14b_480x832x81f", dict(model="wan", double=None, single=None, layers=8), 5.0,
             "8 of 40 blocks at the full 32 760 tokens"),
            ("cogvideox_2b_480x720x49f", dict(model="cogvideox", double=None, single=None, layers=None), 1.0,
             "all 30 blocks")]
    out = {}
    for name, kw, scale, note in runs:
        if kw["model"] == "cogvideox" and world > 1:
            continue  # 30 heads: data parallel only
        a = types.SimpleNamespace(arm="ours", steps=2, warmup=3, no_checkpoint=False, tokens_scale=1.0, check=False,
                                  optimizer="adamw", **kw)
        import gc

        import torch

        def run_variant(ns):
            """One tools/bench_denoiser.py run. EVERY rank executes it (the iteration contains collectives); only rank 0
            gets a record back. Returns (record or None, error string or None)."""
            rec, err = None, None
            try:
                rec = BD.run(ns, manage_dist=False, emit=False)
            except Exception as e:  # noqa: BLE001
                err = f"{type(e).__name__}: {str(e)[:200]}"
            gc.collect()
            torch.cuda.empty_cache()
            return rec, err

        line, err = run_variant(a)
        # the same iteration with selective checkpointing: attention outputs (O, LSE) kept, not recomputed (b200vt.ckpt).
        # Run by ALL ranks, before any rank-dependent `continue` below (rank 0 alone holds the records).
        l2, err2 = (None, "skipped: the full-recompute run failed") if err is not None else run_variant(
            types.SimpleNamespace(**{**vars(a), "keep_attention": True, "warmup": 2}))
        if int(os.environ.get("RANK", "0")) != 0:
            continue
        if err is not None or line is None:
            out[name] = {"error": err or "no record"}
            continue
        kept = None
        if l2 is not None:
            kept = {"s_per_it_measured": l2["s_per_it"], "s_per_it_full_stack": round(l2["s_per_it"] * scale, 3),
                    "it_per_s_full_stack": round(1.0 / (l2["s_per_it"] * scale), 5), "peak_mem_GB": l2["peak_mem_GB"],
                    "attention_share_of_step": l2.get("attention_share_of_step"),
                    "checkpointing": "per block, selective: the attention forward's outputs stay resident (torch selective "
                                     "activation checkpointing, policy b200vt.ckpt.attention_saving_policy); the backward "
                                     "receives the same O and LSE a recomputed forward would produce"}
        elif err2 is not None:
            kept = {"error": err2}
        out[name] = {"s_per_it_measured": line["s_per_it"], "blocks_measured": note,
                     "s_per_it_full_stack": round(line["s_per_it"] * scale, 3),
                     "it_per_s_full_stack": round(1.0 / (line["s_per_it"] * scale), 5),
                     "full_stack_scaling": f"x{scale:g} (identical blocks; embeddings / final layer < 0.1 % of the FLOPs are not in the stack)",
                     "steps": line["steps"], "warmup": line["warmup"], "parallelism": line["config"]["parallelism"],
                     "iteration": "forward + backward + fused AdamW step on the trainable parameters, per-block activation checkpointing",
                     "attention_share_of_step": line.get("attention_share_of_step"),
                     "attention_tflops_in_step": line.get("attention_tflops_in_step"), "peak_mem_GB": line["peak_mem_GB"]}
        if kept is not None:
            out[name]["attention_outputs_kept"] = kept
    if world == 1:
        # BASELINE config 2: the whole VideoCrafter2 3D-UNet LoRA step (tools/bench_vc2_unet.py), ours and the reference's
        # op sequence on the same weights (data parallel model: measured at 1 GPU)
        import types as _t

        import torch
        import bench_vc2_unet as VU
        rec = {}
        for key, arm, graph in (("ours", "ours", False), ("torch", "torch", False), ("ours_cuda_graph", "ours", True)):
            try:
                r = VU.run(_t.SimpleNamespace(arm=arm, steps=5, warmup=3, no_checkpoint=False, check=False, nchw=False,
                                              graph=graph, frozen_bf16=graph), emit=False)
                rec[key] = {"s_per_it": r["s_per_it"], "it_per_s": r["it_per_s"], "peak_mem_GB": r["peak_mem_GB"],
                            "launch": r["config"]["launch"], "activation_checkpointing": r["config"]["activation_checkpointing"],
                            "activation_layout": r["config"]["activation_layout"],
                            "frozen_weights": r["config"]["frozen_weights"]}
            except Exception as e:  # noqa: BLE001
                rec[key] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            import gc
            gc.collect()
            torch.cuda.empty_cache()
        if "s_per_it" in rec.get("ours", {}) and "s_per_it" in rec.get("torch", {}):
            rec["speedup_vs_reference_op_sequence"] = round(rec["torch"]["s_per_it"] / rec["ours"]["s_per_it"], 3)
        if "s_per_it" in rec.get("ours_cuda_graph", {}) and "s_per_it" in rec.get("torch", {}):
            rec["speedup_cuda_graph_vs_reference_op_sequence"] = round(rec["torch"]["s_per_it"] / rec["ours_cuda_graph"]["s_per_it"], 3)
        rec["iteration"] = ("whole UNet (1.4 B parameters): forward on 2 x 4 x 16 x 40 x 64 latents + 77 x 1024 context, MSE loss, "
                            "backward, fused AdamW on rank-4 LoRA adapters; bf16 autocast. `ours` / `torch`: the reference's "
                            "configuration (eager launches, per-block activation checkpointing, use_checkpoint: true), `torch` = "
                            "the reference's op sequence on the same modules and weights; `ours_cuda_graph`: the whole step "
                            "captured once into a CUDA graph, no checkpointing (fits easily in 180 GB), frozen Conv / Linear weights "
                            "stored in bf16 once (bit-identical under autocast)")
        out["videocrafter2_320x512x16f_lora_b2"] = rec
    return out


def rowwise_summary():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_rowwise as BR
    recs = BR.run(iters=5, with_torch=False, emit=False)
    return [{"kernel": r["kernel"], "shape": r["shape"], "GBps": r["ours_GBps"], "frac_of_hbm_peak": r["frac_of_hbm_peak"],
             "ms": r["ours_ms"]} for r in recs] + [{"peak_GBps": recs[0]["peak_GBps"], "peak_kind": recs[0]["peak_kind"],
                                                    "timing": recs[0]["timing"]}]


def sp_parity_check(torch, dist, dev, world, S_loc, leaves, do, step_device, nh: int = 2):
    """One sequence-parallel step against the UNSHARDED kernels on the gathered tensors, for heads [0, nh) (owned by rank
    0 under Ulysses): outputs, dq/dk/dv of this rank's image rows and the text-row gradients. Every rank takes part in the
    gathers; rank 0 computes the comparison. Returns {"max_rel_err": ..., ...} on rank 0, None elsewhere."""
    import math

    import b200vt.ops as ops
    q, k, v, tq, tk, tv = leaves
    out, grads = step_device()

    def gather_rows(t):
        loc = t[:, :, :nh].detach().contiguous()
        parts = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(parts, loc)
        return torch.cat(parts, dim=1)

    full = [torch.cat([gather_rows(a), b[:, :, :nh].detach()], dim=1).contiguous() for a, b in ((q, tq), (k, tk), (v, tv))]
    do_txt = do[:, S_loc:, :nh].float().contiguous()  # the text rows' output is replicated: their upstream gradients add up
    dist.all_reduce(do_txt)
    do_full = torch.cat([gather_rows(do[:, :S_loc]), do_txt.to(torch.bfloat16)], dim=1).contiguous()
    if dist.get_rank() != 0:
        return None
    L = full[0].shape[1]
    scale = 1.0 / math.sqrt(HEAD_DIM)
    o_ref, lse = ops.attn_fwd(*full, None, None, None, L, L, scale)
    dq_r, dk_r, dv_r = ops.attn_bwd(do_full, *full, o_ref, lse, None, None, None, L, L, scale)
    n_img = L - TXT_TOKENS

    def rel(a, b):
        a, b = a.detach(), b.detach()
        return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))

    errs = {"out_img": rel(out[:, :S_loc, :nh], o_ref[:, :S_loc]), "out_txt": rel(out[:, S_loc:, :nh], o_ref[:, n_img:]),
            "dq": rel(grads[0][:, :, :nh], dq_r[:, :S_loc]), "dk": rel(grads[1][:, :, :nh], dk_r[:, :S_loc]),
            "dv": rel(grads[2][:, :, :nh], dv_r[:, :S_loc]), "dq_txt": rel(grads[3][:, :, :nh], dq_r[:, n_img:]),
            "dk_txt": rel(grads[4][:, :, :nh], dk_r[:, n_img:]), "dv_txt": rel(grads[5][:, :, :nh], dv_r[:, n_img:])}
    worst = max(errs.values())
    return {"max_rel_err": round(worst, 6), "tolerance": 2e-2, "ok": bool(worst <= 2e-2), "heads_checked": nh,
            "per_tensor": {n_: round(e, 6) for n_, e in errs.items()},
            "against": "unsharded b200vt attn_fwd / attn_bwd on the gathered (1, 119056, 2, 128) tensors, rank 0"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import b200vt._lib as L
    import b200vt.functional as Fn
    import b200vt.sp as sp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run for N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 must print exactly one line, but NCCL printf()s its "NCCL version ..." banner to stdout when the first
        # communicator is created: send file descriptor 1 to stderr for that moment (and flush C stdio before restoring it)
        import ctypes
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            try:
                ctypes.CDLL(None).fflush(None)
            except Exception:  # noqa: BLE001
                pass
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        if HEADS % world or IMG_TOKENS % world:
            raise SystemExit(f"world size {world} must divide {HEADS} heads and {IMG_TOKENS} image tokens")
    L.call("vt_init", local)

    # ---- synthetic inputs: every rank draws the full tensors from the same seed, then keeps its shard ------------
    S_loc = IMG_TOKENS // world
    g = torch.Generator(device=dev).manual_seed(SEED)

    def draw(rows):
        return torch.randn(1, rows, HEADS, HEAD_DIM, device=dev, dtype=torch.bfloat16, generator=g)

    if world == 1:
        q, k, v = (draw(SEQ).requires_grad_(True) for _ in range(3))
        do = torch.randn(1, SEQ, HEADS * HEAD_DIM, device=dev, dtype=torch.bfloat16, generator=g)
        cu = torch.tensor([0, SEQ, SEQ], dtype=torch.int32, device=dev)  # all text tokens valid (attenion.py:34-57)
        leaves = (q, k, v)

        def step_device(q=q, k=k, v=v, do=do):
            out = Fn.hunyuan_attention(q, k, v, mode="flash", cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=SEQ,
                                       max_seqlen_kv=SEQ, batch_size=1)
            return out, torch.autograd.grad(out, (q, k, v), do)
    else:
        g_img = torch.Generator(device=dev).manual_seed(SEED + 1 + rank)
        q, k, v = (torch.randn(1, S_loc, HEADS, HEAD_DIM, device=dev, dtype=torch.bfloat16,
                               generator=g_img).requires_grad_(True) for _ in range(3))
        tq, tk, tv = (draw(TXT_TOKENS).requires_grad_(True) for _ in range(3))
        do = torch.randn(1, S_loc + TXT_TOKENS, HEADS, HEAD_DIM, device=dev, dtype=torch.bfloat16, generator=g_img)
        attn = sp.UlyssesAttention()
        leaves = (q, k, v, tq, tk, tv)

        def step_device(q=q, k=k, v=v, tq=tq, tk=tk, tv=tv, do=do):
            out = attn(None, q, k, v, dropout_p=0.0, causal=False, joint_tensor_query=tq, joint_tensor_key=tk,
                       joint_tensor_value=tv, joint_strategy="rear")
            return out, torch.autograd.grad(out, (q, k, v, tq, tk, tv), do)

    step_flops = flops_fwd_bwd(SEQ, SEQ)  # whole job, all ranks
    fused_exchange = world > 1 and sp.fused_exchange_available(q, None)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between a barrier + synchronize on both sides; device time from CUDA events; max over ranks."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput ------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    L.profile_enable(True)
    total_ms = timed(step_device, args.steps)
    prof = {name: L.profile_read(kid) for name, kid in (("attn_fwd", L.K_ATTN_FWD), ("attn_bwd", L.K_ATTN_BWD),
                                                        ("attn_bwd_delta", L.K_ATTN_BWD_DELTA),
                                                        ("attn_bwd_dq_convert", L.K_ATTN_BWD_DQ))}
    L.profile_enable(False)
    clk = clocks.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = step_flops / (ms_per_step * 1e-3) / 1e12

    # ---- N > 1: the sequence-parallel step against the unsharded kernels (2 heads), in the driver's own run -------
    sp_parity = None
    if world > 1 and not args.no_extras:
        sp_parity = sp_parity_check(torch, dist, dev, world, S_loc, leaves, do, step_device)

    # ---- end to end with host buffers ----------------------------------------------------------------------------
    # N = 1: the library's host-buffer entry point (functional.HostAttention): pinned (B, L, H, D) tensors in, pinned
    # results out, head groups pipelined over copy-in / compute / copy-out streams. N > 1: each rank copies its
    # sequence shard in, runs the Ulysses step, copies its results out (sequential).
    e2e_steps = max(1, args.steps)  # consecutive calls pipeline into one another: time as many as the device leg
    if world == 1:
        host_in = [t.detach().reshape(1, SEQ, HEADS, HEAD_DIM).to("cpu").pin_memory() for t in (*leaves, do)]
        host_out = [torch.empty((1, SEQ, HEADS, HEAD_DIM), dtype=torch.bfloat16).pin_memory() for _ in range(4)]
        del q, k, v, do, leaves
        torch.cuda.empty_cache()
        host_attn = Fn.HostAttention(1, SEQ, HEADS, HEAD_DIM, head_groups=args.head_groups)

        def step_e2e():
            host_attn(*host_in, *host_out)
    else:
        # every rank's shard lives in pinned host memory; sp.HostUlyssesAttention pipelines head groups: copy-in of group
        # g+1 || all-to-all + attention + backward of group g || copy-out of group g-1
        host_in = [t.detach().to("cpu").pin_memory() for t in (*leaves, do)]  # q, k, v, tq, tk, tv, dO
        host_out = [torch.empty(t.shape, dtype=torch.bfloat16).pin_memory() for t in (do, q, k, v, tq, tk, tv)]
        del q, k, v, tq, tk, tv, do, leaves
        torch.cuda.empty_cache()
        # head groups of at least 3 heads per rank (a 1-head launch is 931 CTAs = 6.3 waves on 148 SMs: the wave tail and the
        # per-group exchanges cost more than the finer overlap gains — N = 8 with 3 groups: e2e 7058 vs 8562 device-resident);
        # with a single group the copies still hide behind the neighbouring steps' kernels (calls pipeline into one another)
        n_groups = max([g_ for g_ in (6, 4, 3, 2) if HEADS % g_ == 0 and (HEADS // g_) % world == 0
                        and (HEADS // g_) // world >= 3] or [1])
        host_attn = sp.HostUlyssesAttention(S_loc, TXT_TOKENS, HEADS, HEAD_DIM, head_groups=n_groups)
        hq, hk, hv, htq, htk, htv, hdo = host_in

        def step_e2e():
            host_attn(hq, hk, hv, hdo, host_out[0], host_out[1], host_out[2], host_out[3], htq, htk, htv,
                      host_out[4], host_out[5], host_out[6])

    step_e2e()  # allocates buffers; untimed
    e2e_ms = timed(step_e2e, e2e_steps) / e2e_steps
    h2d = sum(h.numel() * h.element_size() for h in host_in)
    d2h = sum(h.numel() * h.element_size() for h in host_out)
    e2e_value = step_flops / (e2e_ms * 1e-3) / 1e12

    del host_attn, host_in, host_out
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    def finish(extras, extras_clk, final=True):
        """Rank 0: build and print THE json line from values that all exist before the extras run (pure Python for N > 1:
        no CUDA call), so the extras watchdog below can still emit the headline if a multi-rank extra ever hangs."""
        # ---- roofline of the dominant kernel (tensor-pipe bound) -----------------------------------------------------
        peaks = measured_peaks()
        heads_loc = HEADS // world
        kern = {}
        for name, (ms, n) in prof.items():
            if n:
                kern[name] = {"avg_ms": ms / n, "launches_per_step": n / args.steps}
        fl = {"attn_fwd": flops_fwd(SEQ, SEQ, heads_loc), "attn_bwd": 2.5 * flops_fwd(SEQ, SEQ, heads_loc)}
        dom = max((n for n in ("attn_fwd", "attn_bwd") if n in kern), key=lambda n: kern[n]["avg_ms"])
        achieved = fl[dom] / (kern[dom]["avg_ms"] * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(f"{dom}@k1_h{heads_loc}")
        traffic_source = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of "
                          "this kernel at this shape (a profiler counter: cannot be measured inside the timed run)"
                          if traffic is not None else "no ncu capture committed for this head count")
        roofline = {
            "bound": "tensor", "kernel": f"{dom}_kernel<128>", "achieved": round(achieved, 1),
            "peak": peaks["sustained"], "unit": UNIT, "frac": round(achieved / peaks["sustained"], 4),
            "peak_kind": "sustained bf16 cuBLAS, " + peaks["source"], "frac_of_burst": round(achieved / peaks["burst"], 4),
            "traffic": traffic, "traffic_source": traffic_source,
            "kernels": {n: {"avg_ms": round(kk["avg_ms"], 3), "launches_per_step": kk["launches_per_step"],
                            **({"tflops": round(fl[n] / (kk["avg_ms"] * 1e-3) / 1e12, 1)} if n in fl else {})}
                        for n, kk in kern.items()},
            "step_frac_of_sustained": round(value / world / peaks["sustained"], 4),
        }
        gpu_launches = int(sum(n for _, n in prof.values()))

        # ---- the other BASELINE.json configurations, attention only (N = 1; context for the reader, not the headline) ----
        others = None
        if world == 1:
            others = other_configs(torch, dev)

        # ---- CPU baseline (N = 1 only): oracle port of the reference's torch path on a bounded sample ----------------
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            lq, lk, heads, _ = cpu_size_sample(target_s=12.0)
            tf, dt = cpu_time_sample(lq, lk, heads, steps=1, warmup=0)
            cpu = {"value": round(tf, 4), "unit": UNIT, "cores": host_threads(), "kind": cpu_attention_fn()[1],
                   "sample": sample_text(lq, lk, heads) + f"; {dt:.1f} s"}

        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "B": 1, "L": SEQ, "img_tokens": IMG_TOKENS, "txt_tokens": TXT_TOKENS,
                       "H": HEADS, "D": HEAD_DIM, "flops_per_step": step_flops,
                       "parallelism": "single" if world == 1 else f"ulysses_sp{world}",
                       **({} if world == 1 else {"exchange": (
                           "q/k/v + gradients: NCCL all_to_all; forward O: peer stores from the attention epilogue "
                           "(symmetric memory over NVLink)" if fused_exchange else "NCCL all_to_all")}),
                       "l2": "inputs (4 x 731 MB) exceed the 126 MB L2; no flush needed"},
            "clocks": clk,
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps,
                    "api": ("functional.HostAttention (head-group pipeline)" if world == 1
                            else "sp.HostUlyssesAttention (head-group pipeline: copy-in || all-to-all + attention + backward || copy-out)")},
            "gpu_launches": gpu_launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            **({"other_configs": others} if others else {}),
            **({"sp_parity": sp_parity} if sp_parity is not None else {}),
            **extras,
            **({"extras_clocks": extras_clk} if extras_clk is not None else {}),
        }
        print(json.dumps(line), flush=True)
        if final and world > 1:
            dist.barrier()
            dist.destroy_process_group()

    # ---- the metric's second half and the claims around the headline, measured in this run (all ranks take part) -----
    extras, extras_clk = {}, None
    watchdog = None
    if world > 1 and not args.no_extras:
        # The extras at N > 1 are multi-rank iterations full of collectives: if one of them ever hangs, rank 0 still prints
        # the complete headline line (without the extras) and every rank exits 0, instead of the whole run timing out.
        limit_s = float(os.environ.get("B200VT_BENCH_EXTRAS_TIMEOUT_S", "480"))
        note = {"denoiser_it_s": {"error": f"the multi-rank extras did not finish within {limit_s:.0f} s (watchdog); the "
                                           "headline keys of this line are complete"}}
        watchdog = start_extras_watchdog(rank, limit_s, lambda: finish(note, None, final=False))
    if not args.no_extras:
        xclocks = ClockSampler(local)
        if rank == 0:
            xclocks.start()
        try:
            extras["denoiser_it_s"] = denoiser_it_s(world)
        except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
            extras["denoiser_it_s"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        if world == 1:
            for key, fn in (("library_baselines", lambda: library_baselines(torch, dev)), ("rowwise", rowwise_summary)):
                try:
                    extras[key] = fn()
                except Exception as e:  # noqa: BLE001
                    extras[key] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
                gc.collect()
                torch.cuda.empty_cache()
        extras_clk = xclocks.stop() if rank == 0 else None
    if watchdog is not None:
        watchdog.cancel()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    finish(extras, extras_clk)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~15 s host-core leg (profiling runs)")
    ap.add_argument("--head-groups", type=int, default=4, help="pipeline depth of the host-buffer (e2e) entry point")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip sp_parity / library_baselines / denoiser_it_s / rowwise (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: --warmup {args.warmup} is below the 3 the timing rules ask for", file=sys.stderr)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"),
               os.path.abspath(__file__), *sys.argv[1:]]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
