"""ISP (Intern Sequence Parallel) weight-parallel communicator.

Every ``ISPLinear`` keeps a ``1/wp`` row-shard of its weight; the full weight is all-gathered over the WEIGHT group just
in time for the forward and again for the backward GEMMs, and the weight gradient is reduce-scattered (AVG) back to the
shard owner (reference ``internlm/core/communication/isp.py:31-567``).

What is kept from the reference: module ordering per block, prefetch of the *next* module's weight while the current
one computes (previous module in backward), a small rotating memory pool for gathered weights so the allocator is not
hit every layer, deferred accumulation of reduce-scattered gradients, the scheduler hook that flips forward/backward
state.  What changes: gathers are single ``all_gather_into_tensor`` calls into pooled buffers (no list API, no copies),
gradient results are accumulated straight into the optimizer's persistent gradient arena.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch import nn

from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import SchedulerHook


class ISPCommModelConfig:
    """dtype / device / activation-checkpoint fraction of the model (reference ``isp.py:31-43``)."""

    def __init__(self, dtype: torch.dtype = torch.half, device: torch.device = None, activation_checkpointing: float = 0.0,
                 module_shapes: Dict[str, torch.Size] = None) -> None:
        self.dtype = dtype
        self.device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
                                 else torch.device("cpu"))
        self.activation_checkpointing = activation_checkpointing
        self.module_shapes = module_shapes


class MemoryPool:
    """Rotating pool of gathered-weight buffers keyed by shape (reference ``isp.py:45-140``): with prefetch there are at
    most two gathered weights of a given shape alive (current + next), so ``depth`` = 3 is always safe."""

    def __init__(self, dtype, device, depth: int = 3, module_shapes: Optional[Dict[str, torch.Size]] = None):
        self.dtype, self.device, self.depth = dtype, device, depth
        self.module_shapes = module_shapes
        self._bufs: Dict[Tuple[int, ...], List[torch.Tensor]] = {}
        self._next: Dict[Tuple[int, ...], int] = {}
        self._named: Dict[tuple, torch.Tensor] = {}
        self._rs: Dict[tuple, List[torch.Tensor]] = {}
        self._rs_busy: Dict[tuple, List[bool]] = {}

    def get(self, shape) -> torch.Tensor:
        key = tuple(shape)
        if key not in self._bufs:
            self._bufs[key] = [torch.empty(key, dtype=self.dtype, device=self.device) for _ in range(self.depth)]
            self._next[key] = 0
        i = self._next[key]
        self._next[key] = (i + 1) % self.depth
        return self._bufs[key][i]

    def reset_lazy_pools(self):
        """Give every lent reduce-scatter buffer back (``train.py`` calls it once per step)."""
        for key in self._rs_busy:
            self._rs_busy[key] = [False] * len(self._rs_busy[key])

    # ---- the reference's block-indexed interface (``isp.py:73-122``) on the same storage --------------------------------
    def allocate_constant_zero(self, size: tuple) -> torch.Tensor:
        """One shared all-zero tensor per shape (stands in for a gradient that arrives later through ``flush_grads``)."""
        key = ("zero", tuple(size))
        if key not in self._named:
            self._named[key] = torch.zeros(tuple(size), dtype=self.dtype, device=self.device)
        return self._named[key]

    def allocate_all_gather_memory(self, block_index: int, module_name: str, is_bias: bool = False, shape=None) -> torch.Tensor:
        """Gathered-weight buffer of ``module_name`` in block ``block_index``: two generations (``block_index % 2``) are alive
        at a time - the block that computes and the one being prefetched.  ``shape`` is needed on first use unless the pool was
        given ``module_shapes``."""
        key = ("ag", block_index % 2, module_name, bool(is_bias))
        if key not in self._named:
            shape = shape if shape is not None else (self.module_shapes or {}).get(module_name)
            assert shape is not None, f"no shape known for {module_name!r}"
            shape = tuple(shape)[:1] if is_bias else tuple(shape)
            self._named[key] = torch.empty(shape, dtype=self.dtype, device=self.device)
        return self._named[key]

    def allocate_reduce_scatter_memory(self, key: tuple) -> torch.Tensor:
        """A free buffer of shape ``key`` (grown on demand); the tensor remembers its slot in ``index_in_pool``."""
        key = tuple(key)
        bufs, busy = self._rs.setdefault(key, []), self._rs_busy.setdefault(key, [])
        for i, b in enumerate(busy):
            if not b:
                busy[i] = True
                return bufs[i]
        t = torch.zeros(key, dtype=self.dtype, device=self.device)
        t.index_in_pool = len(bufs)
        bufs.append(t)
        busy.append(True)
        return t

    def free_reduce_scatter_memory(self, key, index) -> None:
        self._rs_busy[tuple(key)][index] = False


class ISPOverlapState:
    """Structure of one model chunk as the communicator sees it (reference ``isp.py:124-140``): its transformer blocks, the
    weight-parallel linears inside each block in execution order, the ones outside any block (the head), how many leading
    blocks are activation-checkpointed (their forward runs a second time during backward, so their weights are gathered in
    forward order again), and the gathers currently in flight keyed by module."""

    def __init__(self) -> None:
        self.num_blocks: int = 0
        self.ckpt_block_num: int = 0
        self.isp_outs: List[nn.Module] = []
        self.isp_modules: List[nn.Module] = []
        self.index_to_isp_modules: Dict[int, List[nn.Module]] = {}
        self.index_to_block: Dict[int, nn.Module] = {}
        self.module_to_index: Dict[nn.Module, int] = {}
        self.weight_global_handle: Dict[nn.Module, Optional[dist.Work]] = {}
        self.weight_global_output: Dict[nn.Module, torch.Tensor] = {}
        self.bias_global_handle: Dict[nn.Module, Optional[dist.Work]] = {}
        self.bias_global_output: Dict[nn.Module, torch.Tensor] = {}

    def block_of(self, module: nn.Module) -> Optional[int]:
        return self.module_to_index.get(module)

    def in_flight(self) -> int:
        return len(self.weight_global_output) + len(self.bias_global_output)


def _parse_chunk(chunk: nn.Module, isp_cls, checkpoint_fraction: float) -> ISPOverlapState:
    st = ISPOverlapState()
    blocks: List[nn.Module] = []
    for child in chunk.modules():
        if isinstance(child, nn.ModuleList) and len(child) > 0 and all(
                any(isinstance(m, isp_cls) for m in b.modules()) for b in child):
            blocks = list(child)
            break
    st.num_blocks = len(blocks)
    st.ckpt_block_num = int(checkpoint_fraction * st.num_blocks)
    inside = set()
    for i, block in enumerate(blocks):
        st.index_to_block[i] = block
        mods = [m for m in block.modules() if isinstance(m, isp_cls)]
        st.index_to_isp_modules[i] = mods
        for m in mods:
            st.module_to_index[m] = i
            inside.add(id(m))
        st.isp_modules += mods
    st.isp_outs = [m for m in chunk.modules() if isinstance(m, isp_cls) and id(m) not in inside]
    st.isp_modules += st.isp_outs
    return st


class ISPCommunicator:
    def __init__(self, model: Union[nn.Module, nn.ModuleList], model_conf: ISPCommModelConfig, overlap: bool = False,
                 enable_memory_pool: bool = False, process_group: dist.ProcessGroup = None) -> None:
        from internevo_b200.parallel.linear import ISPLinear

        self.process_group = process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.overlap = overlap
        self.model_conf = model_conf
        self.is_forward = True
        self.memory_pool = MemoryPool(model_conf.dtype, model_conf.device,
                                      module_shapes=model_conf.module_shapes) if enable_memory_pool else None
        self._prerequisites: list = []
        chunks = list(model) if isinstance(model, nn.ModuleList) else [model]
        chunks = [getattr(c, "model", c) for c in chunks]
        self._order: List[nn.Module] = []
        for c in chunks:
            mods = c if isinstance(c, nn.ModuleList) else [c]
            for m in mods:
                self._order += [sub for sub in m.modules() if isinstance(sub, ISPLinear)]
        self._index = {id(m): i for i, m in enumerate(self._order)}
        # per-chunk structure (blocks, modules per block, checkpointed blocks); the interleaved scheduler names the chunk
        # that is about to run through ``switch_current_model_chunk``
        self._overlap_states: Dict[int, ISPOverlapState] = {}
        for cid, c in enumerate(chunks):
            self._overlap_states[cid] = _parse_chunk(c, ISPLinear, float(model_conf.activation_checkpointing or 0.0))
        self._cur_chunk = 0
        self._gathered: Dict[int, Tuple[torch.Tensor, Optional[dist.Work]]] = {}
        self._pending_grads: List[Tuple[torch.nn.Parameter, torch.Tensor, Optional[dist.Work]]] = []

    # ---------------------------------------------------------------------------------------------------------------
    @property
    def overlap_state(self) -> ISPOverlapState:
        """State of the chunk that is currently executing."""
        return self._overlap_states[self._cur_chunk]

    def switch_current_model_chunk(self, chunk_id: int) -> None:
        """Interleaved pipeline: the scheduler announces which model chunk runs next (reference ``isp.py:431-442``)."""
        assert chunk_id in self._overlap_states, (chunk_id, list(self._overlap_states))
        self._cur_chunk = chunk_id

    def _state_of(self, module) -> ISPOverlapState:
        for st in self._overlap_states.values():
            if module in st.module_to_index or module in st.isp_outs:
                return st
        return self.overlap_state

    def _launch_gather(self, module, weight: torch.Tensor, async_op: bool):
        shape = (weight.shape[0] * self.world, *weight.shape[1:])
        out = self.memory_pool.get(shape) if self.memory_pool is not None else torch.empty(
            shape, dtype=weight.dtype, device=weight.device)
        h = dist.all_gather_into_tensor(out, weight.contiguous(), group=self.process_group, async_op=async_op)
        return out, (h if async_op else None)

    def all_gather_weight(self, module, weight: torch.Tensor, is_bias: bool = False, backward: bool = False):
        """Full weight (or bias) of ``module``; consumes a prefetched gather when one is in flight and starts the
        prefetch of the neighbouring module (next in forward, previous in backward)."""
        if self.world <= 1:
            return weight
        if not backward:
            for fn in self._prerequisites:     # e.g. "this module's parameter broadcast has landed"
                fn(module)
        if is_bias:
            out, _ = self._launch_gather(module, weight, async_op=False)
            return out
        key = id(module)
        if key in self._gathered:
            out, h = self._gathered.pop(key)
            st = self._state_of(module)
            st.weight_global_output.pop(module, None)
            st.weight_global_handle.pop(module, None)
            if h is not None:
                h.wait()
        else:
            out, _ = self._launch_gather(module, weight, async_op=False)
        if self.overlap:
            i = self._index.get(key)
            if i is not None:
                j = i - 1 if backward else i + 1
                if 0 <= j < len(self._order):
                    nxt = self._order[j]
                    fused_fwd, fused_bwd = getattr(nxt, "_b200_isp_fused", (False, False))
                    if id(nxt) not in self._gathered and not (fused_bwd if backward else fused_fwd):
                        self._gathered[id(nxt)] = self._launch_gather(nxt, nxt.weight, async_op=True)
                        st = self._state_of(nxt)
                        st.weight_global_output[nxt], st.weight_global_handle[nxt] = self._gathered[id(nxt)]
        return out

    # ---- the reference's entry points (``isp.py:444-526``) ------------------------------------------------------------------
    def register_prerequisite_for_forward_prefetch_hooks(self, prerequisite_func) -> None:
        """``prerequisite_func(module)`` runs before a module's weight is gathered in forward (the parameter-broadcast handler
        registers its wait here so a gather never reads a shard the optimizer is still updating)."""
        self._prerequisites.append(prerequisite_func)

    def all_gather(self, tensor: torch.Tensor, module: nn.Module, is_bias: bool = False):
        return self.all_gather_weight(module, tensor, is_bias=is_bias, backward=not self.is_forward)

    def reduce_scatter(self, tensor: torch.Tensor, module: nn.Module, op=dist.ReduceOp.AVG, is_bias: bool = False):
        """``(gradient shard or None, None)``: with overlap the shard is folded into the gradient arena by ``flush_grads`` and
        autograd gets ``None``."""
        assert op == dist.ReduceOp.AVG, "weight gradients are averaged over the weight-parallel group"
        param = module.bias if is_bias else module.weight
        return self.reduce_scatter_grad(module, param, tensor, is_bias=is_bias), None

    def release_weight(self, module):
        """Gathered buffers come from the rotating pool (or the caching allocator): nothing to free explicitly."""

    def reduce_scatter_grad(self, module, param, grad_full: torch.Tensor, is_bias: bool = False):
        """AVG reduce-scatter of the full-weight gradient to this rank's shard. With overlap the result is accumulated
        into the parameter's gradient buffer later (``flush_grads``) and ``None`` is returned to autograd."""
        if self.world <= 1:
            return grad_full
        out = torch.empty(grad_full.shape[0] // self.world, *grad_full.shape[1:], dtype=grad_full.dtype,
                          device=grad_full.device)
        if grad_full.is_cuda:
            h = dist.reduce_scatter_tensor(out, grad_full.contiguous(), op=dist.ReduceOp.AVG, group=self.process_group,
                                           async_op=self.overlap)
        else:
            h = dist.reduce_scatter_tensor(out, grad_full.contiguous(), group=self.process_group, async_op=self.overlap)
        if not self.overlap:
            if not grad_full.is_cuda:
                out.div_(self.world)
            return out
        self._pending_grads.append((param, out, h, grad_full))
        return None

    def flush_grads(self):
        """Wait for outstanding reduce-scatters and fold them into the gradient arena (or ``.grad``)."""
        for param, out, h, keep in self._pending_grads:
            if h is not None:
                h.wait()
            if not out.is_cuda:
                out.div_(self.world)
            buf = getattr(param, "grad_buf", None)
            if buf is not None:
                if getattr(param, "grad_ready", False):
                    buf.add_(out)
                else:
                    buf.copy_(out)
                    param.grad_ready = True
            elif param.grad is None:
                param.grad = out
            else:
                param.grad.add_(out)
            del keep
        self._pending_grads = []


class ISPCommunicatorSchedulerHook(SchedulerHook):
    """Flips the communicator between forward and backward mode and flushes deferred gradients (reference
    ``isp.py:529-567``)."""

    def __init__(self, overlap_handler: ISPCommunicator, zero_optim) -> None:
        self._isp_communicator = overlap_handler
        self._zero_optim = zero_optim

    def before_forward(self, scheduler, inputs) -> None:
        self._isp_communicator.is_forward = True
        gpc.is_forward = True

    def after_forward(self, scheduler, outputs) -> None:
        pass

    def before_criterion(self, scheduler, outputs, label) -> None:
        pass

    def after_criterion(self, scheduler, loss) -> None:
        pass

    def before_backward(self, scheduler, outputs, outputs_grad) -> None:
        self._isp_communicator.is_forward = False
        gpc.is_forward = False

    def after_backward(self, scheduler, inputs_grad) -> None:
        self._isp_communicator.flush_grads()
        self._isp_communicator._gathered.clear()
        for st in self._isp_communicator._overlap_states.values():
            st.weight_global_output.clear()
            st.weight_global_handle.clear()

    def post_helper_func(self, scheduler, outputs, label) -> None:
        pass
