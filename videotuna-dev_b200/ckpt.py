"""Activation checkpointing that keeps the attention outputs.

The reference checkpoints whole transformer blocks — `torch.utils.checkpoint.checkpoint(ckpt_wrapper(block), *args,
use_reentrant=False)` (hunyuan/hyvideo_i2v/modules/models.py:764-769, 789-794) — so the backward pass re-runs every block's
forward, the joint attention included: at HunyuanVideo's 119 056 tokens that is one more 140 ms attention forward per
block, 17 % of the finetuning iteration, to save 742 MB (O: L x H x D bf16, plus the log-sum-exp) per block. On a B200 the
memory is there (60 blocks: 44.6 GB of 180 GB), so the B200-native choice is to keep exactly those two tensors and recompute
the rest: PyTorch's selective activation checkpointing with a policy that marks the b200vt attention forward ops MUST_SAVE.
The backward kernels receive exactly the O and LSE a recomputed forward would produce, so the gradients are those of full
recomputation (up to the arrival order of dQ's fp32 reduce-adds, which varies from run to run either way).

    from b200vt import ckpt
    y = ckpt.checkpoint(block, *args)                       # = torch's checkpoint(use_reentrant=False, context_fn=...)
    ckpt.keep_attention_in_checkpoints()                    # or: make every non-reentrant torch checkpoint call do it
                                                            # (the reference's own call sites, unmodified)

The policy acts on the REGISTERED ops (`torch.ops.b200vt.*`): under the checkpoint's dispatch mode `ops._eager_ok` is false, so
the library's eager fast path steps aside by itself.
"""
from __future__ import annotations

import functools
from typing import Callable

import torch
import torch.utils.checkpoint as _tuc

from . import ops  # noqa: F401  (registers torch.ops.b200vt.*)

#: ops whose outputs stay resident between the forward and the backward of a checkpointed region
SAVED_OPS = ("attn_fwd", "ulysses_attn_fwd")


def _saved_overloads():
    out = set()
    for name in SAVED_OPS:
        pkt = getattr(torch.ops.b200vt, name, None)
        if pkt is not None:
            out.add(pkt.default)
    return out


def attention_saving_policy(ctx, op, *args, **kwargs):
    """Selective-checkpoint policy: keep the attention forward's outputs (O, log-sum-exp; for the sequence-parallel op also
    the exchanged sequence-layout output), recompute everything else."""
    if op in _saved_overloads():
        return _tuc.CheckpointPolicy.MUST_SAVE
    return _tuc.CheckpointPolicy.PREFER_RECOMPUTE


def context_fn():
    """`context_fn` for torch.utils.checkpoint.checkpoint(..., use_reentrant=False)."""
    saved = _saved_overloads()  # resolved once per checkpointed region, not once per op

    def policy(ctx, op, *args, **kwargs):
        return _tuc.CheckpointPolicy.MUST_SAVE if op in saved else _tuc.CheckpointPolicy.PREFER_RECOMPUTE

    return _tuc.create_selective_checkpoint_contexts(policy)


def checkpoint(fn: Callable, *args, **kwargs):
    """torch.utils.checkpoint.checkpoint(fn, *args, use_reentrant=False) that does not recompute attention."""
    kwargs.setdefault("use_reentrant", False)
    if kwargs["use_reentrant"]:
        raise ValueError("selective checkpointing needs use_reentrant=False")
    kwargs.setdefault("context_fn", context_fn)
    return _ORIGINAL(fn, *args, **kwargs)


_ORIGINAL = _tuc.checkpoint


def keep_attention_in_checkpoints(enable: bool = True) -> None:
    """Make every `torch.utils.checkpoint.checkpoint(..., use_reentrant=False)` call that passes no `context_fn` of its own use
    the attention-saving policy — the reference's call sites look the function up on the module at call time
    (`torch.utils.checkpoint.checkpoint(...)`), so they pick this up unmodified. Re-entrant calls and calls with their own
    context_fn are passed through untouched. `enable=False` restores torch's function."""
    if not enable:
        _tuc.checkpoint = _ORIGINAL
        return

    @functools.wraps(_ORIGINAL)
    def wrapped(function, *args, use_reentrant=None, context_fn=None, **kwargs):
        if use_reentrant is False and context_fn is None:
            context_fn = globals()["context_fn"]
        if context_fn is not None:
            kwargs["context_fn"] = context_fn
        return _ORIGINAL(function, *args, use_reentrant=use_reentrant, **kwargs)

    wrapped._b200vt_wrapped = True
    _tuc.checkpoint = wrapped
