"""Symmetric peer-memory heap over NVLink / NVSwitch.

``torch.distributed._symmetric_memory`` is used ONLY for the plumbing (cuMem allocation, handle exchange over the
process group's store, peer mapping); it hands back one base pointer per rank.  Everything that moves data through those
pointers — barriers, reduce-scatter, all-gather, the GEMM-fused collectives — is our own kernel code
(``csrc/comm_kernels.cu``, ``csrc/gemm_sm100.cu``).

A ``SymmBuffer`` is one symmetric allocation (same size on every rank of the group) plus device-resident pointer tables;
``SymmFlags`` is a symmetric ``uint32`` array with a monotonically increasing epoch: a flag is "set" when it holds a
value ``>=`` the epoch of the operation, so flags never need to be cleared between operations.
"""
from __future__ import annotations

import os
import socket
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

_BARRIER_SLOTS = 1024        # [0, 1024): barrier flags (one slot per peer)
_AG_BASE = 1024              # [1024, 8192): all-gather chunk flags (local)
_RS_BASE = 8192              # [8192, ...): reduce-scatter tile flags, `world` blocks of tiles each
_DONE_BASE = 512             # [512, 512 + world): end-of-kernel completion words of the fused all-reduce GEMM
AG_FLAG_WORDS = _RS_BASE - _AG_BASE
RS_FLAG_WORDS = 8 * 8192
FLAG_WORDS = 8192 + RS_FLAG_WORDS


# B200_SYMM_DEBUG=1: race / protocol debugging of the in-kernel peer-memory signalling.  Every back-end
# poisons the slab it is about to receive into with NaN (after a barrier, so no in-flight reader is disturbed), and
# asserts after the exchange that nothing it consumes is still poison: a missing flag wait, a wrong slab address or a
# too-early read shows up as a hard error instead of a silently wrong gradient.  Costs a host sync per exchange.
DEBUG = os.environ.get("B200_SYMM_DEBUG", "0") == "1"


def poison(t: torch.Tensor) -> None:
    t.fill_(float("nan"))


def assert_clean(t: torch.Tensor, what: str) -> None:
    bad = int(torch.isnan(t.float()).sum())
    if bad:
        raise RuntimeError(f"[B200_SYMM_DEBUG] {what}: {bad} of {t.numel()} consumed elements were never written by a peer")


def symm_available() -> bool:
    if not torch.cuda.is_available():
        return False
    try:
        import importlib

        importlib.import_module("torch.distributed._symmetric_memory")
        return True
    except Exception:
        return False


# ---- which groups are peer-addressable -------------------------------------------------------------------------------------
# Peer loads / stores reach the GPUs of ONE NVLink domain - one node here.  A process group that spans nodes (data parallel over
# two 8-GPU boxes, a pipeline across nodes) keeps its collectives on NCCL; groups inside a node (tensor, weight, expert, a ZeRO
# sub-group) get the peer-memory kernels.  Every rank publishes an identity of its node in the rendezvous store once, right
# after ``init_process_group`` (host-side key / value traffic only - no communicator is created for it), and a group is
# peer-addressable when all its members published the same identity.
_node_ids: Optional[List[str]] = None


def this_node() -> str:
    """Identity of the machine (and container) this rank runs on.  ``B200_NODE_ID`` overrides it (tests simulate several nodes
    on one box with it; a site whose containers share one IPC namespace can give them one id)."""
    forced = os.environ.get("B200_NODE_ID")
    if forced:
        return forced
    try:
        with open("/proc/sys/kernel/random/boot_id") as f:
            boot = f.read().strip()
    except OSError:
        boot = ""
    return f"{socket.gethostname()}/{boot}"


def exchange_node_ids() -> Optional[List[str]]:
    """Publish this rank's node identity and read everybody's (called by ``ParallelContext.init_global_dist``)."""
    global _node_ids
    _node_ids = None
    if not dist.is_initialized():
        return None
    rank, world = dist.get_rank(), dist.get_world_size()
    try:
        store = dist.distributed_c10d._get_default_store()
        store.set(f"b200/node_id/{rank}", this_node())
        _node_ids = [store.get(f"b200/node_id/{r}").decode() for r in range(world)]
    except Exception:  # pragma: no cover - a launcher without a key / value store: fall back to the launcher's own statement
        local = int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0)
        _node_ids = ["node0"] * world if local == world else None
    return _node_ids


def node_ids() -> Optional[List[str]]:
    return _node_ids


def group_is_intra_node(group: Optional[dist.ProcessGroup]) -> bool:
    """True when every rank of ``group`` runs on this rank's node.  Pure table look-up after ``exchange_node_ids``; a process
    that never exchanged (the groups were made outside ``gpc``) asks the group itself once."""
    if group is None or not dist.is_initialized():
        return True
    if _node_ids is not None:
        ranks = dist.get_process_group_ranks(group)
        return len({_node_ids[r] for r in ranks}) == 1
    key = id(group)
    if key not in _intra_cache:
        seen = [None] * dist.get_world_size(group)
        dist.all_gather_object(seen, this_node(), group=group)
        _intra_cache[key] = len(set(seen)) == 1
    return _intra_cache[key]


_intra_cache: Dict[int, bool] = {}


def peer_addressable(group: Optional[dist.ProcessGroup]) -> bool:
    """Can the peer-memory kernels serve ``group``: symmetric memory is there and the group does not leave the node."""
    return symm_available() and group_is_intra_node(group)


class SymmBuffer:
    """A symmetric allocation of ``numel`` elements of ``dtype`` on every rank of ``group``."""

    def __init__(self, numel: int, dtype: torch.dtype, group: dist.ProcessGroup, zero: bool = True):
        import torch.distributed._symmetric_memory as symm_mem

        from internevo_b200.ops import _lib

        assert _lib.available(), "peer-memory kernels need the native extension"

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.tensor = symm_mem.empty(numel, dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()))
        if zero:
            self.tensor.zero_()
        self.handle = symm_mem.rendezvous(self.tensor, group)
        self.base_ptrs: List[int] = [int(p) for p in self.handle.buffer_ptrs]
        # NVSwitch multicast (NVLS) mapping of the same allocation: 0 when the fabric / driver has no multicast support
        try:
            self.mc_ptr: int = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        except Exception:  # pragma: no cover - older torch builds raise instead of returning 0
            self.mc_ptr = 0
        self.elem = self.tensor.element_size()
        self._tables: Dict[int, torch.Tensor] = {}
        dist.barrier(group)

    def ptr_table(self, elem_offset: int = 0) -> torch.Tensor:
        """Device ``int64[world]`` with every rank's pointer to element ``elem_offset`` of this buffer."""
        t = self._tables.get(elem_offset)
        if t is None:
            t = torch.tensor([p + elem_offset * self.elem for p in self.base_ptrs], dtype=torch.int64,
                             device=self.tensor.device)
            self._tables[elem_offset] = t
        return t

    def table_ptr(self, elem_offset: int = 0) -> int:
        return self.ptr_table(elem_offset).data_ptr()

    def view(self, elem_offset: int, shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        return self.tensor[elem_offset: elem_offset + n].view(*shape)


class SymmFlags(SymmBuffer):
    """Symmetric flag words + the group's epoch counter and a stream-ordered device barrier."""

    def __init__(self, group: dist.ProcessGroup, words: int = FLAG_WORDS):
        super().__init__(words, torch.int32, group, zero=True)
        self.epoch = 0
        torch.cuda.synchronize()
        dist.barrier(group)

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def barrier(self):
        """All ranks of the group rendezvous on the current stream (no host involvement)."""
        torch.ops.b200.symm_barrier(self.table_ptr(0), self.rank, self.world, self.next_epoch())


_flags_cache: Dict[int, SymmFlags] = {}


def flags_for(group: dist.ProcessGroup) -> SymmFlags:
    key = id(group)
    if key not in _flags_cache:
        _flags_cache[key] = SymmFlags(group)
    return _flags_cache[key]


def ag_flag_table(flags: SymmFlags) -> int:
    return flags.table_ptr(_AG_BASE)


def rs_flag_table(flags: SymmFlags) -> int:
    return flags.table_ptr(_RS_BASE)


def done_flag_table(flags: SymmFlags) -> int:
    return flags.table_ptr(_DONE_BASE)
