"""A stand-in for the few xfuser names the reference's sequence-parallel entry points import.

The reference enters sequence parallelism through xfuser [ext, 0.4.3.post2, not vendored]:
    wan      usp_attn_forward / usp_dit_forward    videotuna/models/wan/wan/distributed/xdit_context_parallel.py:3-7
             bound at                              videotuna/models/wan/wan/text2video.py:261-271, flow/wanvideo.py:120-131
    hunyuan  parallelize_transformer               videotuna/flow/hunyuanvideo.py:31-46, 114-178, 303-318
and cannot even import those modules without the package. `install()` (called by patch.patch_sp()) registers this module
under the names `xfuser`, `xfuser.core`, `xfuser.core.distributed`, `xfuser.core.distributed.parallel_state` and
`xfuser.core.long_ctx_attention` — ONLY when the real package is absent — so the unmodified reference flows run with
`xFuserLongContextAttention` = b200vt.sp.UlyssesAttention (all-to-all over NCCL/NVLink + the sm_100a attention kernels).

Scope: Ulysses only (ring_degree must be 1, as in every reference config: flow/hunyuanvideo.py:194-196 defaults,
configs/008_wanvideo); the sequence-parallel group is the whole world or consecutive-rank subgroups of it.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import sp

_STATE = {"group": None, "degree": None}
#: Attention core used by xFuserLongContextAttention(); None = the CUDA kernels. The CPU (gloo) tests inject the oracle here.
ATTN_FN: Optional[Callable] = None


def init_distributed_environment(rank: int = -1, world_size: int = -1, local_rank: int = -1, backend: Optional[str] = None,
                                 **_unused) -> None:
    """xfuser.core.distributed.init_distributed_environment: make sure torch.distributed is up (env:// rendezvous)."""
    if dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if rank >= 0 and world_size > 0:
        kw = dict(rank=rank, world_size=world_size)
    dist.init_process_group(backend, **kw)


def initialize_model_parallel(sequence_parallel_degree: Optional[int] = None, ring_degree: int = 1,
                              ulysses_degree: Optional[int] = None, **_unused) -> None:
    """xfuser.core.distributed.initialize_model_parallel for the degrees the reference passes (flow/wanvideo.py:126-130,
    flow/hunyuanvideo.py:315-319): consecutive ranks form one Ulysses group."""
    if ring_degree not in (None, 1):
        raise NotImplementedError("b200vt sequence parallelism is Ulysses-only (ring_degree must be 1)")
    world = dist.get_world_size() if dist.is_initialized() else 1
    degree = sequence_parallel_degree or ulysses_degree or world
    if ulysses_degree not in (None, degree):
        raise ValueError(f"ulysses_degree {ulysses_degree} != sequence_parallel_degree {degree} with ring_degree 1")
    if world % degree != 0:
        raise ValueError(f"sequence_parallel_degree {degree} does not divide the world size {world}")
    _STATE["degree"] = degree
    if degree == world or world == 1:
        _STATE["group"] = None  # the default (WORLD) group
        return
    mine, rank = None, dist.get_rank()
    for first in range(0, world, degree):  # every rank must create every group, in the same order
        ranks = list(range(first, first + degree))
        g = dist.new_group(ranks)
        if rank in ranks:
            mine = g
    _STATE["group"] = mine


def get_sequence_parallel_world_size() -> int:
    return sp._world(_STATE["group"])


def get_sequence_parallel_rank() -> int:
    return sp._rank(_STATE["group"])


class _GroupCoordinator:
    """What the reference uses of xfuser's GroupCoordinator: `get_sp_group().all_gather(x, dim=...)`
    (xdit_context_parallel.py:142, flow/hunyuanvideo.py:173) plus the size / rank attributes."""

    @property
    def device_group(self):
        return _STATE["group"]

    @property
    def world_size(self) -> int:
        return get_sequence_parallel_world_size()

    @property
    def rank_in_group(self) -> int:
        return get_sequence_parallel_rank()

    def all_gather(self, input_: torch.Tensor, dim: int = -1, separate_tensors: bool = False):
        if separate_tensors:
            P = self.world_size
            parts = [torch.empty_like(input_) for _ in range(P)]
            if P == 1:
                return [input_]
            dist.all_gather(parts, input_.contiguous(), group=_STATE["group"])
            return parts
        return sp.gather_sequence(input_, dim=dim, group=_STATE["group"])


_SP_GROUP = _GroupCoordinator()


def get_sp_group() -> _GroupCoordinator:
    return _SP_GROUP


def get_world_group() -> _GroupCoordinator:
    return _SP_GROUP


class xFuserLongContextAttention(sp.UlyssesAttention):
    """`xFuserLongContextAttention()` as the reference constructs it (no arguments; xdit_context_parallel.py:179,
    flow/hunyuanvideo.py:157, attenion.py:169): Ulysses attention over the group initialize_model_parallel() made."""

    def __init__(self, *_args, **_kwargs):
        super().__init__(group=_STATE["group"], attn_fn=ATTN_FN)


_NAMES = ("init_distributed_environment", "initialize_model_parallel", "get_sequence_parallel_world_size",
          "get_sequence_parallel_rank", "get_sp_group", "get_world_group")


def install(force: bool = False) -> bool:
    """Register the stand-in modules unless the real xfuser is importable. Returns True when the shim is (now) active."""
    if isinstance(sys.modules.get("xfuser"), types.ModuleType) and getattr(sys.modules["xfuser"], "_b200vt_shim", False):
        return True
    if not force:
        try:
            if importlib.util.find_spec("xfuser") is not None:
                return False
        except (ImportError, ValueError):
            pass
    me = sys.modules[__name__]
    root, core = types.ModuleType("xfuser"), types.ModuleType("xfuser.core")
    distributed, state = types.ModuleType("xfuser.core.distributed"), types.ModuleType("xfuser.core.distributed.parallel_state")
    lca = types.ModuleType("xfuser.core.long_ctx_attention")
    for m in (root, core, distributed, state, lca):
        m._b200vt_shim = True
        m.__path__ = []  # importable as packages
    for n in _NAMES:
        setattr(distributed, n, getattr(me, n))
        setattr(state, n, getattr(me, n))
    lca.xFuserLongContextAttention = xFuserLongContextAttention
    root.core, core.distributed, core.long_ctx_attention, distributed.parallel_state = core, distributed, lca, state
    sys.modules.update({"xfuser": root, "xfuser.core": core, "xfuser.core.distributed": distributed,
                        "xfuser.core.distributed.parallel_state": state, "xfuser.core.long_ctx_attention": lca})
    return True


def uninstall() -> None:
    for n in [k for k, m in sys.modules.items() if k.split(".")[0] == "xfuser" and getattr(m, "_b200vt_shim", False)]:
        del sys.modules[n]
    _STATE.update(group=None, degree=None)
