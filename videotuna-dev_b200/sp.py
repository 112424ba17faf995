"""Ulysses sequence parallelism for the DiT attention path (one process per GPU, NCCL all-to-all over NVLink/NVSwitch).

Replaces, for the hot path only, what the reference delegates to xfuser's `xFuserLongContextAttention` [ext, xfuser
0.4.3.post2; not vendored, parity unpinned]:
    hunyuan  parallel_attention(hybrid_seq_parallel_attn, ...)     videotuna/models/hunyuan/hyvideo_t2v/modules/attenion.py:159-212
    hunyuan  parallelize_transformer (sets block.hybrid_seq_parallel_attn)  videotuna/flow/hunyuanvideo.py:114-178
    wan      usp_attn_forward                                       videotuna/models/wan/wan/distributed/xdit_context_parallel.py:149-192
and follows the autograd semantics of the reference's only in-tree Ulysses code (OpenSora `_AllToAll`,
videotuna/models/opensora/acceleration/communications.py:23-58): the backward of an all-to-all is the all-to-all with
scatter/gather dims swapped.

Layout: activations live sequence-sharded (B, L/P, H, D). One all-to-all turns q, k, v into (B, L, H/P, D); the local
kernel attends over the full sequence for H/P heads; one all-to-all returns (B, L/P, H, D). With B == 1 (every
BASELINE config) the receive side of seq->head and the send side of head->seq are plain views, so each exchange costs
exactly one pack copy. Replicated text tokens ("rear" joint strategy) are head-sliced locally and written behind the
gathered image tokens in the same buffer, so no concatenation copy exists either.

SP correctness is *defined* as equality with the single-GPU result on the same inputs.
"""
from __future__ import annotations

import importlib
import math
import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def _rank(group) -> int:
    return dist.get_rank(group) if dist.is_initialized() else 0


def _a2a(send: Tensor, group, out: Optional[Tensor] = None) -> Tensor:
    """all_to_all_single over dim 0 (P equal chunks). `send` must be contiguous; `out` (optional) likewise."""
    recv = torch.empty_like(send) if out is None else out
    dist.all_to_all_single(recv, send, group=group)
    return recv


def _pack_seq_to_head(x: Tensor, P: int) -> Tensor:
    """(B, Ls, H, D) -> send buffer (P, B, Ls, H/P, D): chunk p carries head group p of the local tokens."""
    B, Ls, H, D = x.shape
    return x.reshape(B, Ls, P, H // P, D).permute(2, 0, 1, 3, 4).contiguous()


def _unpack_seq_to_head(recv: Tensor) -> Tensor:
    """receive buffer (P, B, Ls, Hp, D), index 0 = source rank = sequence chunk -> (B, P*Ls, Hp, D)."""
    P, B, Ls, Hp, D = recv.shape
    if B == 1:
        return recv.view(1, P * Ls, Hp, D)
    return recv.permute(1, 0, 2, 3, 4).reshape(B, P * Ls, Hp, D)


def _pack_head_to_seq(y: Tensor, P: int) -> Tensor:
    """(B, L, Hp, D) -> send buffer (P, B, L/P, Hp, D): chunk p carries sequence chunk p of the local head group."""
    B, L, Hp, D = y.shape
    if B == 1 and y.is_contiguous():
        return y.view(P, 1, L // P, Hp, D)
    return y.reshape(B, P, L // P, Hp, D).permute(1, 0, 2, 3, 4).contiguous()


def _unpack_head_to_seq(recv: Tensor) -> Tensor:
    """receive buffer (P, B, Ls, Hp, D), index 0 = source rank = head group -> (B, Ls, P*Hp, D)."""
    P, B, Ls, Hp, D = recv.shape
    return recv.permute(1, 2, 0, 3, 4).reshape(B, Ls, P * Hp, D)


class _SeqToHead(torch.autograd.Function):
    """(B, L/P, H, D) sequence-sharded -> (B, L + T, H/P, D) head-sharded, with T optional rows left free at the rear
    (filled by the caller with the head slice of replicated text tokens)."""

    @staticmethod
    def forward(ctx, x: Tensor, group, rear: int):
        P = _world(group)
        ctx.group, ctx.P, ctx.rear = group, P, rear
        B, Ls, H, D = x.shape
        if H % P != 0:
            raise ValueError(f"Ulysses needs num_heads % world_size == 0 (H={H}, P={P})")
        Hp = H // P
        send = _pack_seq_to_head(x, P)
        if B == 1:
            full = x.new_empty((1, P * Ls + rear, Hp, D))
            _a2a(send, group, out=full[:, : P * Ls].view(P, 1, Ls, Hp, D))
            return full
        out = _unpack_seq_to_head(_a2a(send, group))
        if rear:
            out = torch.cat([out, out.new_empty((B, rear, Hp, D))], dim=1)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        P = ctx.P
        L = g.shape[1] - ctx.rear
        recv = _a2a(_pack_head_to_seq(g[:, :L], P), ctx.group)
        return _unpack_head_to_seq(recv), None, None


class _HeadToSeq(torch.autograd.Function):
    """(B, L, H/P, D) head-sharded -> (B, L/P, H, D) sequence-sharded."""

    @staticmethod
    def forward(ctx, y: Tensor, group):
        P = _world(group)
        ctx.group, ctx.P = group, P
        if y.shape[1] % P != 0:
            raise ValueError(f"sequence length {y.shape[1]} is not divisible by the SP world size {P}")
        recv = _a2a(_pack_head_to_seq(y, P), group)
        return _unpack_head_to_seq(recv)

    @staticmethod
    def backward(ctx, g: Tensor):
        recv = _a2a(_pack_seq_to_head(g, ctx.P), ctx.group)
        return _unpack_seq_to_head(recv), None


class _GatherHeads(torch.autograd.Function):
    """(B, T, H/P, D) per-rank head group of the replicated text rows -> (B, T, H, D) on every rank.
    Adjoint: every rank's gradient for my head group is summed (reduce-scatter written as all-to-all + sum)."""

    @staticmethod
    def forward(ctx, t: Tensor, group):
        P = _world(group)
        ctx.group, ctx.P = group, P
        parts = [torch.empty_like(t) for _ in range(P)]
        dist.all_gather(parts, t.contiguous(), group=group)
        return torch.cat(parts, dim=2)

    @staticmethod
    def backward(ctx, g: Tensor):
        P = ctx.P
        B, T, H, D = g.shape
        send = g.reshape(B, T, P, H // P, D).permute(2, 0, 1, 3, 4).contiguous()
        recv = _a2a(send, ctx.group)
        return recv.sum(dim=0), None


def seq_to_head(x: Tensor, group=None, rear: int = 0) -> Tensor:
    if _world(group) == 1:
        return x if rear == 0 else torch.cat([x, x.new_empty((x.shape[0], rear, *x.shape[2:]))], dim=1)
    return _SeqToHead.apply(x, group, rear)


def head_to_seq(y: Tensor, group=None) -> Tensor:
    if _world(group) == 1:
        return y
    return _HeadToSeq.apply(y, group)


class _FillRear(torch.autograd.Function):
    """Write `joint` (B, T, Hp, D) into the last T rows of `full` in place (no concatenation copy)."""

    @staticmethod
    def forward(ctx, full: Tensor, joint: Tensor):
        T = joint.shape[1]
        ctx.T = T
        full[:, full.shape[1] - T:].copy_(joint)
        ctx.mark_dirty(full)
        return full

    @staticmethod
    def backward(ctx, g: Tensor):
        T = ctx.T
        L = g.shape[1] - T
        return g, g[:, L:]


# ---------------------------------------------------------------------------------------------------------------------
# fused exchange: the attention kernel's epilogue writes O straight into the destination ranks' buffers (NVLink peer
# stores through symmetric memory), replacing the output all-to-all and its pack copy
# ---------------------------------------------------------------------------------------------------------------------
class _PeerBuffers:
    """One symmetric (S_loc, H, D) bf16 output buffer per (group, shape), with every rank's peer-mapped address."""
    _cache: dict = {}

    def __init__(self, group, s_loc: int, H: int, D: int, device):
        import torch.distributed._symmetric_memory as symm
        pg = group if group is not None else dist.group.WORLD
        self.buf = symm.empty((s_loc, H, D), dtype=torch.bfloat16, device=device)
        self.handle = symm.rendezvous(self.buf, pg)
        self.ptrs = [int(a) for a in self.handle.buffer_ptrs]

    @classmethod
    def get(cls, group, s_loc: int, H: int, D: int, device) -> "_PeerBuffers":
        key = (id(group), s_loc, H, D, str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(group, s_loc, H, D, device)
        return cls._cache[key]


_GROUPS: dict = {}  # id(group) -> group: process groups cannot travel through an op schema


@torch.library.custom_op("b200vt::ulysses_attn_fwd", mutates_args=(), device_types="cuda")
def ulysses_attn_fwd(q: Tensor, k: Tensor, v: Tensor, softmax_scale: float, n_img: int,
                     group_key: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Head-sharded attention with the head -> sequence exchange of its output fused into the kernel epilogue, as ONE op:
    symmetric-memory barrier, `attn_fwd_scatter` (peer stores over NVLink), barrier, copy of the received rows out of the
    symmetric buffer. Returns (out_seq (1, L/P, H, D), o (1, L [+ T], H/P, D) head layout, lse). One op so that selective
    activation checkpointing (ckpt.py) can keep all three and skip the barriers, the kernel and the copy on every rank alike."""
    from . import ops
    group = _GROUPS[group_key]
    P, r = _world(group), _rank(group)
    _, Ltot, Hp, D = q.shape
    s_loc = n_img // P
    pb = _PeerBuffers.get(group, s_loc, Hp * P, D, q.device)
    ptrs = [a + r * Hp * D * 2 for a in pb.ptrs]  # this rank's head slot inside every destination row
    pb.handle.barrier(channel=0)                   # every rank has consumed the previous contents of its buffer
    o, lse = ops.attn_fwd_scatter(q, k, v, None, float(softmax_scale), pb.buf, ptrs, s_loc, Hp * P * D, D)
    pb.handle.barrier(channel=1)                   # all peers' stores into this rank's buffer are complete
    return pb.buf.unsqueeze(0).clone(), o, lse


@ulysses_attn_fwd.register_fake
def _(q, k, v, softmax_scale, n_img, group_key):
    _, Ltot, Hp, D = q.shape
    P = _world(_GROUPS[group_key])
    return (q.new_empty((1, n_img // P, Hp * P, D)), q.new_empty((1, Ltot, Hp, D)),
            q.new_empty((1, Hp, Ltot), dtype=torch.float32))


class _FusedAttnExchange(torch.autograd.Function):
    """out_seq (1, L/P, H, D) = head_to_seq(attention(q, k, v)) for head-sharded q, k, v (1, L [+ T], H/P, D), with the
    exchange done by the kernel epilogue. Backward: the adjoint exchange of dO (NCCL all-to-all, `seq_to_head`) followed
    by the ordinary backward kernel on the locally kept head-layout O. Rows past L (replicated text) come back as a
    second, local tensor."""

    @staticmethod
    def forward(ctx, q, k, v, group, scale, n_img):
        key = id(group)
        _GROUPS.setdefault(key, group)
        out_seq, o, lse = ulysses_attn_fwd(q, k, v, float(scale), int(n_img), key)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.group, ctx.scale, ctx.n_img = group, scale, n_img
        return out_seq, o[:, n_img:]

    @staticmethod
    def backward(ctx, d_seq, d_txt):
        from . import ops
        q, k, v, o, lse = ctx.saved_tensors
        P = _world(ctx.group)
        d_head = _unpack_seq_to_head(_a2a(_pack_seq_to_head(d_seq.contiguous(), P), ctx.group))  # (1, L, H/P, D)
        if o.shape[1] > ctx.n_img:
            d_txt = d_txt if d_txt is not None else torch.zeros_like(o[:, ctx.n_img:])
            d_head = torch.cat([d_head, d_txt], dim=1)
        dq, dk, dv = ops.attn_bwd(d_head, q, k, v, o, lse, None, None, None, q.shape[1], k.shape[1], float(ctx.scale))
        return dq, dk, dv, None, None, None


def fused_exchange_available(q: Tensor, group) -> bool:
    """The fused epilogue needs CUDA bf16, head dim 128, batch 1, P <= 8 and torch symmetric memory over NVLink."""
    if not (q.is_cuda and q.dtype == torch.bfloat16 and q.shape[0] == 1 and q.shape[-1] == 128):
        return False
    if os.environ.get("B200VT_SP_FUSED", "1") == "0" or not dist.is_initialized():
        return False
    P = _world(group)
    if P < 2 or P > 8:
        return False
    try:
        importlib.import_module("torch.distributed._symmetric_memory")
    except Exception:  # noqa: BLE001
        return False
    return dist.get_backend(group) == "nccl"


def _default_attn(q: Tensor, k: Tensor, v: Tensor, softmax_scale: Optional[float]) -> Tensor:
    from . import functional as Fn  # CUDA kernels; raises without the library — there is no CPU path
    return Fn.attention_blhd(q, k, v, softmax_scale=softmax_scale)


class UlyssesAttention:
    """Drop-in for the object the reference stores in `block.hybrid_seq_parallel_attn` (flow/hunyuanvideo.py:154-157)
    and calls as `xFuserLongContextAttention()(None, q, k, v, ...)`.

    __call__(attn, query, key, value, *, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1,-1),
             joint_tensor_query=None, joint_tensor_key=None, joint_tensor_value=None, joint_strategy="none")
    query/key/value: (B, L/P, H, D) local shard. joint_*: (B, T, H, D) replicated on every rank; with
    joint_strategy="rear" they are attended jointly behind the image tokens. Returns (B, L/P [+ T], H, D).
    """

    def __init__(self, group=None, attn_fn: Optional[Callable] = None):
        self.group = group
        self.attn_fn = _default_attn if attn_fn is None else attn_fn

    def __call__(self, attn, query: Tensor, key: Tensor, value: Tensor, *, dropout_p: float = 0.0,
                 softmax_scale: Optional[float] = None, causal: bool = False, window_size=(-1, -1),
                 joint_tensor_query: Optional[Tensor] = None, joint_tensor_key: Optional[Tensor] = None,
                 joint_tensor_value: Optional[Tensor] = None, joint_strategy: str = "none") -> Tensor:
        if causal or dropout_p != 0.0 or tuple(window_size) != (-1, -1):
            raise NotImplementedError("causal / dropout / windowed attention are not on the sequence-parallel path")
        has_joint = joint_tensor_query is not None
        if has_joint and joint_strategy != "rear":
            raise NotImplementedError(f"joint_strategy={joint_strategy!r}: only 'rear' is used by the reference")
        P, r = _world(self.group), _rank(self.group)
        H = query.shape[2]
        if H % P != 0:
            # Head counts that do not divide the group size (CogVideoX-2B: 30 heads on 4 or 8 GPUs): pad with zero heads up
            # to the next multiple — a zero head attends uniformly over zero values, contributes nothing, costs
            # (pad / H) extra attention FLOPs (2 of 32 there) — and drop them from the result; gradients flow through the pad.
            pad = P - H % P

            def padh(t):
                return None if t is None else torch.nn.functional.pad(t, (0, 0, 0, pad))

            out = self(attn, padh(query), padh(key), padh(value), dropout_p=dropout_p, softmax_scale=softmax_scale,
                       causal=causal, window_size=window_size, joint_tensor_query=padh(joint_tensor_query),
                       joint_tensor_key=padh(joint_tensor_key), joint_tensor_value=padh(joint_tensor_value),
                       joint_strategy=joint_strategy)
            return out[:, :, :H]
        Hp = H // P
        T = joint_tensor_query.shape[1] if has_joint else 0
        q = seq_to_head(query, self.group, rear=T)
        k = seq_to_head(key, self.group, rear=T)
        v = seq_to_head(value, self.group, rear=T)
        if T:
            sl = slice(r * Hp, (r + 1) * Hp)
            q = _FillRear.apply(q, joint_tensor_query[:, :, sl])
            k = _FillRear.apply(k, joint_tensor_key[:, :, sl])
            v = _FillRear.apply(v, joint_tensor_value[:, :, sl])
        if self.attn_fn is _default_attn and fused_exchange_available(q, self.group):
            # attention with the output exchange fused into its epilogue (peer stores over NVLink)
            scale = 1.0 / math.sqrt(q.shape[-1]) if softmax_scale is None else float(softmax_scale)
            img, txt = _FusedAttnExchange.apply(q, k, v, self.group, scale, q.shape[1] - T)
            if not T:
                return img
            if P > 1:
                txt = _GatherHeads.apply(txt, self.group)
            return torch.cat([img, txt], dim=1)
        out = self.attn_fn(q, k, v, softmax_scale)
        if not T:
            return head_to_seq(out, self.group)
        L = out.shape[1] - T
        img = head_to_seq(out[:, :L], self.group)
        txt = out[:, L:]
        if P > 1:
            txt = _GatherHeads.apply(txt, self.group)
        return torch.cat([img, txt], dim=1)


class HostUlyssesAttention:
    """Sequence-parallel joint attention forward + backward for activations that live in PINNED HOST memory (the
    activation-offload case of long-sequence DiT finetuning, one rank's shard per process): the multi-GPU counterpart of
    functional.HostAttention. Heads are independent, so the H heads are cut into groups of `hg` (a multiple of the SP
    world size) and pipelined on three streams: group g+1's q, k, v, dO shards are copied in (strided DMA out of the
    (1, L/P, H, D) host layout) while group g runs all-to-all -> attention -> all-to-all -> backward, and group g-1's
    out, dq, dk, dv are copied out. The replicated text tensors (a few MB) travel once per call.
    Same arithmetic as UlyssesAttention + autograd on device tensors (bench.py sp_parity / tests)."""

    def __init__(self, s_loc: int, T: int, H: int, D: int, head_groups: int, group=None, device=None):
        P = _world(group)
        if H % head_groups != 0 or (H // head_groups) % P != 0:
            raise ValueError(f"head_groups={head_groups} must divide H={H} into groups that are multiples of the SP size {P}")
        self.s_loc, self.T, self.H, self.D, self.G, self.hg = s_loc, T, H, D, head_groups, H // head_groups
        self.group = group
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.attn = UlyssesAttention(group)

        def buf(rows):
            return torch.empty((1, rows, self.hg, D), dtype=torch.bfloat16, device=self.dev)

        self.inp = [[buf(s_loc), buf(s_loc), buf(s_loc), buf(s_loc + T)] for _ in range(2)]
        self.txt = [[torch.empty((1, T, H, D), dtype=torch.bfloat16, device=self.dev) for _ in range(3)] for _ in range(2)] if T else None
        self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.ev_in = [torch.cuda.Event() for _ in range(2)]      # inputs of slot landed
        self.ev_free = [torch.cuda.Event() for _ in range(2)]    # compute finished reading the slot's inputs
        self.ev_done = [torch.cuda.Event() for _ in range(2)]    # outputs of slot computed
        self.ev_out = [torch.cuda.Event() for _ in range(2)]     # outputs of slot copied out
        self.ev_txt_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_txt_free = [torch.cuda.Event() for _ in range(2)]
        self.outp = [None, None]
        self.n_groups_done = 0   # running counters: slots alternate ACROSS calls, so that the copy-in of the next call's
        self.n_calls = 0         # first group overlaps the compute of this call's last group

    def synchronize(self) -> None:
        """Wait until every copy-out enqueued so far has landed in the host buffers."""
        self.s_out.synchronize()

    def __call__(self, q: Tensor, k: Tensor, v: Tensor, dout: Tensor, out: Tensor, dq: Tensor, dk: Tensor, dv: Tensor,
                 tq: Optional[Tensor] = None, tk: Optional[Tensor] = None, tv: Optional[Tensor] = None,
                 dtq: Optional[Tensor] = None, dtk: Optional[Tensor] = None, dtv: Optional[Tensor] = None):
        """q, k, v, dq, dk, dv: pinned host (1, L/P, H, D); dout, out: (1, L/P + T, H, D); tq/tk/tv and their gradient
        buffers: (1, T, H, D) (replicated text, "rear"). A rank's text gradients cover its own head slices only (zeros
        elsewhere), as with UlyssesAttention. Everything is ENQUEUED and consecutive calls pipeline into one another (the
        next call's copy-in runs under this call's kernels, this call's last copy-out under the next call's): the input
        buffers must stay untouched, and the outputs are valid, only after synchronize() (or a device synchronize)."""
        from .functional import copy_head_group
        cur = torch.cuda.current_stream(self.dev)
        T = self.T
        first = self.n_calls == 0
        if first:  # nothing in flight yet: order the side streams behind whatever produced the host buffers' device peers
            self.s_in.wait_stream(cur)
            self.s_out.wait_stream(cur)
        tslot = self.n_calls & 1
        if T:
            with torch.cuda.stream(self.s_in):
                if self.n_calls >= 2:
                    self.s_in.wait_event(self.ev_txt_free[tslot])
                for d_t, h_t in zip(self.txt[tslot], (tq, tk, tv)):
                    d_t.copy_(h_t, non_blocking=True)
                self.ev_txt_in[tslot].record(self.s_in)
            t_grads = [torch.zeros_like(t) for t in self.txt[tslot]]
        for g in range(self.G):
            n = self.n_groups_done
            slot, h0 = n & 1, g * self.hg
            with torch.cuda.stream(self.s_in):
                if n >= 2:
                    self.s_in.wait_event(self.ev_free[slot])  # the compute that last used this slot's inputs is done
                for d_t, h_t in zip(self.inp[slot], (q, k, v, dout)):
                    copy_head_group(d_t, h_t, h0, True, self.s_in)
                self.ev_in[slot].record(self.s_in)
            cur.wait_event(self.ev_in[slot])
            if g == 0 and T:
                cur.wait_event(self.ev_txt_in[tslot])
            qd, kd, vd, dod = self.inp[slot]
            leaves = [t.detach().requires_grad_(True) for t in (qd, kd, vd)]
            kw = {}
            if T:
                tl = [t[:, :, h0:h0 + self.hg].detach().requires_grad_(True) for t in self.txt[tslot]]
                kw = dict(joint_tensor_query=tl[0], joint_tensor_key=tl[1], joint_tensor_value=tl[2], joint_strategy="rear")
                leaves += tl
            o = self.attn(None, *leaves[:3], **kw)
            grads = torch.autograd.grad(o, leaves, dod)
            if T:
                for acc, gr in zip(t_grads, grads[3:]):
                    acc[:, :, h0:h0 + self.hg].copy_(gr)
            self.ev_free[slot].record(cur)
            self.ev_done[slot].record(cur)
            results = (o.detach(), *grads[:3])
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_done[slot])
                for d_t, h_t in zip(results, (out, dq, dk, dv)):
                    d_t.record_stream(self.s_out)  # freed by the caching allocator only after the copy-out has run
                    copy_head_group(d_t, h_t, h0, False, self.s_out)
            self.n_groups_done += 1
        if T:
            self.ev_txt_free[tslot].record(cur)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                for acc, h_t in zip(t_grads, (dtq, dtk, dtv)):
                    if h_t is not None:
                        acc.record_stream(self.s_out)
                        h_t.copy_(acc, non_blocking=True)
        self.n_calls += 1
        return out, dq, dk, dv


def ulysses_attention(q: Tensor, k: Tensor, v: Tensor, group=None, softmax_scale: Optional[float] = None,
                      attn_fn: Optional[Callable] = None) -> Tensor:
    """Plain Ulysses attention on sequence-sharded (B, L/P, H, D) tensors (Wan usp_attn_forward core)."""
    return UlyssesAttention(group, attn_fn)(None, q, k, v, softmax_scale=softmax_scale)


def shard_sequence(x: Tensor, dim: int = 1, group=None) -> Tensor:
    """torch.chunk along `dim` and keep this rank's piece (wan usp_dit_forward, xdit_context_parallel.py:121-127)."""
    P, r = _world(group), _rank(group)
    if x.shape[dim] % P != 0:
        raise ValueError(f"dim {dim} of size {x.shape[dim]} is not divisible by the SP world size {P}")
    return x.chunk(P, dim=dim)[r]


def gather_sequence(x: Tensor, dim: int = 1, group=None) -> Tensor:
    """get_sp_group().all_gather(x, dim) at the end of the denoiser (xdit_context_parallel.py:142,
    flow/hunyuanvideo.py:173). Forward only, like the reference (inference path)."""
    P = _world(group)
    if P == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(P)]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, dim=dim)
