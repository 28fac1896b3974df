"""1F1B and interleaved-1F1B pipeline schedules (reference ``internlm/core/scheduler/pipeline_scheduler.py:111-1430``).

Schedule shape is the reference's (warm-up forwards, steady one-forward-one-backward, cool-down backwards; virtual
chunks for the interleaved variant with ``micro_num % pp == 0``).  Differences: activations are static-shape 2-D
tensors whose shape is computed once (no per-step meta handshake unless the caller disables it), p2p never
host-synchronises, and the interleaved overlap starts the exchange *before* the compute it hides and collects it after
(plain handles instead of generator coroutines).
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Callable, List, Optional, Tuple, Union

import torch
import torch.distributed as dist

import internevo_b200.core.communication as comm
from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.core.naive_amp import NaiveAMPModel
from internevo_b200.utils.common import SchedulerHook, get_current_device, move_to_device
from internevo_b200.utils.timeout import llm_timeout

from .base_scheduler import BaseScheduler, model_kwargs


@contextmanager
def switch_virtual_pipeline_parallel_rank(rank):
    """Temporarily make ``rank`` the current model chunk (reference ``pipeline_scheduler.py:91-98``)."""
    prev = gpc.virtual_pipeline_parallel_rank
    try:
        gpc.set_virtual_pipeline_parallel_rank(rank)
        yield
    finally:
        gpc.set_virtual_pipeline_parallel_rank(prev)


@contextmanager
def switch_optimizer_grad_sync_skip_mode(optimizer, skip: bool = True):
    """Temporarily (un)set the optimizer's "do not reduce gradients yet" flag (reference ``:101-108``)."""
    prev = optimizer.skip_grad_reduce
    try:
        optimizer.skip_grad_reduce = skip
        yield
    finally:
        optimizer.skip_grad_reduce = prev


def get_tensor_shape():
    """Activation shape crossing a stage boundary: ``[tokens (/sp), hidden]`` (reference ``:31-69``)."""
    if gpc.config is not None and gpc.config.get("TENSOR_SHAPE", None) is not None:
        return tuple(gpc.config.TENSOR_SHAPE)      # explicit override for user models whose boundary tensor is not [tokens, hidden]
    if not gpc.is_initialized(ParallelMode.PIPELINE) or gpc.config is None:
        return None
    data, model = gpc.config.get("data", None), gpc.config.get("model", None)
    if not data or not model or "hidden_size" not in model:
        return None
    # packed or not, one micro-batch carries micro_bsz * seq_len tokens; the stages exchange them flattened
    tokens = data["micro_bsz"] * data["seq_len"]
    sp = gpc.get_world_size(ParallelMode.TENSOR) if gpc.config.parallel.get("sequence_parallel", False) else 1
    return (tokens // sp, model["hidden_size"])


def pack_return_tensors(return_tensors):
    output, label = tuple(zip(*return_tensors))
    if isinstance(output[0], torch.Tensor):
        output = torch.cat(output, dim=0)
    elif isinstance(output[0], (list, tuple)):
        output = tuple(torch.cat(t, dim=0) for t in zip(*output))
    else:
        raise TypeError("Output of model must be tensor or list/tuple of tensors")
    if isinstance(label[0], torch.Tensor):
        label = torch.cat(label, dim=0)
    else:
        merged = {k: [] for k in label[0].keys()}
        for d in label:
            for k, v in d.items():
                merged[k].append(v)
        label = {k: torch.cat(v, dim=0) for k, v in merged.items()}
    return output, label


class PipelineScheduler(BaseScheduler):
    """Non-interleaved 1F1B."""

    def __init__(self, num_microbatches: int, dtype: torch.dtype = torch.float, data_process_func: Callable = None,
                 tensor_shape: Union[torch.Size, List[int], Tuple[int]] = None, scatter_gather_tensors: bool = False,
                 scheduler_hooks: Optional[List[SchedulerHook]] = None):
        assert num_microbatches > 0, f"expected num_microbatches > 0, got {num_microbatches}"
        assert not isinstance(tensor_shape, int), "tensor_shape must be a sequence"
        super().__init__(data_process_func=data_process_func)
        self.num_microbatches = num_microbatches
        self.dtype = dtype
        self._hooks = scheduler_hooks or []
        self._tensor_shape = tuple(tensor_shape) if tensor_shape is not None else None
        self.scatter_gather_tensors = (
            scatter_gather_tensors and gpc.is_initialized(ParallelMode.TENSOR)
            and gpc.get_world_size(ParallelMode.TENSOR) > 1 and not gpc.config.parallel.get("sequence_parallel", False)
        )
        self.batch_data = self.batch_label = None
        self.microbatch_offset = 0
        self.bsz_stride = 1
        self.batch_size = None

    @property
    def tensor_shape(self):
        return self._tensor_shape

    @tensor_shape.setter
    def tensor_shape(self, tensor_shape):
        self._tensor_shape = tuple(tensor_shape) if tensor_shape is not None else None

    def pre_processing(self, engine):
        pass

    def _call_hooks(self, func_name: str, *args, **kwargs) -> None:
        for hook in self._hooks:
            getattr(hook, func_name)(self, *args, **kwargs)

    # ------------------------------------------------------------------------------------------------------------
    def load_batch(self, engine, data_iter):
        batch_data, self.batch_size = engine.load_batch(data_iter, to_gpu=False)
        assert self.batch_size % self.num_microbatches == 0, "Batch size should divided by the number of microbatches"
        self.bsz_stride = self.batch_size // self.num_microbatches
        self.microbatch_offset = 0
        self.batch_data, self.batch_label = batch_data

    def load_micro_batch(self):
        data, label = self._load_micro_batch(self.batch_data, self.batch_label, self.microbatch_offset, self.bsz_stride)
        if self.data_process_func:
            data["input_ids"] = self.data_process_func(data["input_ids"], data["cu_seqlens"])
            label = self.data_process_func(label, data["cu_seqlens"], **self._label_pad)
            data.pop("cu_seqlens")
            data.pop("indexes")
            data.pop("max_seqlen", None)
        self.microbatch_offset += self.bsz_stride
        ms = data.pop("max_seqlen", None)
        data, label = move_to_device(data), move_to_device(label)
        if ms is not None:
            data["max_seqlen"] = ms
        return data, label

    def _get_data_label_for_current_step(self, stage_output, micro_batch_data):
        if isinstance(micro_batch_data, (tuple, list)):
            if gpc.is_first_rank(ParallelMode.PIPELINE):
                data, label = micro_batch_data[:2]
                return data, label
            return stage_output, micro_batch_data[1]
        return micro_batch_data, None

    @staticmethod
    def _prep_stage_input(input_obj, data: dict) -> dict:
        """Non-first stages consume the received activation instead of ``input_ids``.  The activation crosses the stage boundary
        flattened to ``[tokens, hidden]``; for an UN-packed batch (validation, ``use_packed_dataset=False``: ``input_ids`` is
        ``[rows, seq]`` and there is no ``cu_seqlens``) the row structure would be lost with it - the later stage would attend
        across rows and count positions through them - so it is handed over as uniform segments, exactly what the first
        stage derives from ``input_ids`` itself."""
        if input_obj is not None:
            data = dict(data)
            data["hidden_states"] = input_obj
            ids = data.pop("input_ids", None)
            if torch.is_tensor(ids) and ids.dim() == 2 and data.get("cu_seqlens", None) is None:
                rows, seq = ids.shape
                data["cu_seqlens"] = torch.arange(0, (rows + 1) * seq, seq, device=input_obj.device, dtype=torch.int32)
                data["indexes"] = torch.arange(seq, device=input_obj.device).repeat(rows)
                data["max_seqlen"] = seq
        return data

    def _call_engine(self, engine, data):  # pylint: disable=W0237
        if data is None:
            return None
        return engine(**model_kwargs(engine, data)) if isinstance(data, dict) else engine(data)

    def _forward_step(self, engine, input_obj, return_tensors, return_output_label=True, accum_loss=None,
                      accum_moe_loss=None, model=None):
        """One micro-batch forward on this stage; last stage also evaluates the loss (scaled by 1/M)."""
        data, label = self.load_micro_batch()
        data = self._prep_stage_input(input_obj, data)
        is_moe = hasattr(gpc.config.model, "num_experts")
        self._call_hooks("before_forward", data)
        run = engine if model is None else model
        out = run(**model_kwargs(run, data))
        if is_moe:
            output_obj, moe_losses = out
        else:
            output_obj, moe_losses = out, None
        self._call_hooks("after_forward", output_obj)
        if gpc.is_pipeline_last_stage():
            self._call_hooks("post_helper_func", output_obj, label)
            if return_output_label:
                return_tensors.append((output_obj, label))
            if accum_loss is not None:
                self._call_hooks("before_criterion", output_obj, label)
                loss = self._call_engine_criterion(engine, output_obj, label)
                self._call_hooks("after_criterion", loss)
                loss_reduced = loss / self.num_microbatches
                accum_loss.add_(loss_reduced.detach())
                output_obj = loss_reduced
        moe_loss = None
        if is_moe and gpc.config.model.num_experts > 1 and moe_losses:
            moe_loss = sum(moe_losses) * gpc.config.loss.moe_loss_coeff / self.num_microbatches
            if accum_moe_loss is not None:
                accum_moe_loss.add_(moe_loss.detach())
        return output_obj, moe_loss

    def _backward_step(self, engine, step_id, input_obj, output_obj, output_obj_grad, moe_loss=None):
        """Backward of one micro-batch on this stage; returns the gradient w.r.t. the stage input."""
        if input_obj is not None:
            if isinstance(input_obj, torch.Tensor):
                input_obj.retain_grad()
            else:
                for t in input_obj:
                    if t is not None:
                        t.retain_grad()
        if engine.optimizer is not None:
            engine.optimizer.skip_grad_reduce = step_id != self.num_microbatches - 1
        self._call_hooks("before_backward", output_obj, output_obj_grad)
        if moe_loss is None or not moe_loss.requires_grad:
            if output_obj_grad is None:
                engine.backward(output_obj)
            else:
                engine.backward_by_grad(output_obj, output_obj_grad)
        else:
            # chain rule: this stage's auxiliary loss is a second root of the same graph (reference ``:363-377``)
            scale = engine.optimizer.grad_scaler.scale if hasattr(engine.optimizer, "grad_scaler") else 1.0
            if output_obj_grad is None:
                engine.backward(output_obj + moe_loss)
            else:
                engine.backward_by_grad([output_obj, moe_loss * scale], [output_obj_grad, None])
        self._call_hooks("after_backward", None)
        if input_obj is None:
            return None
        if isinstance(input_obj, torch.Tensor):
            return input_obj.grad
        return [t.grad if t is not None else None for t in input_obj]

    # ------------------------------------------------------------------------------------------------------------
    def _shape_handshake(self, output_obj, need_send: bool):
        if need_send and self._tensor_shape is None:
            comm.send_obj_meta(output_obj)

    def _forward_only_step(self, engine, return_loss=True, return_output_label=True):
        return_tensors = []
        accum_loss = torch.zeros(1, device=get_current_device()) if return_loss and gpc.is_pipeline_last_stage(True) else None
        accum_moe_loss = torch.zeros(1, device=get_current_device())
        ft_shape = self._tensor_shape
        for _ in range(self.num_microbatches):
            if not gpc.is_first_rank(ParallelMode.PIPELINE):
                if ft_shape is None:
                    ft_shape = comm.recv_obj_meta()
                input_obj = comm.recv_forward(ft_shape, dtype=self.dtype, scatter_gather_tensors=self.scatter_gather_tensors)
            else:
                input_obj = None
            output_obj, _ = self._forward_step(engine, input_obj, return_tensors, return_output_label, accum_loss,
                                               accum_moe_loss)
            if not gpc.is_last_rank(ParallelMode.PIPELINE):
                if self._tensor_shape is None:
                    comm.send_obj_meta(output_obj)
                comm.send_forward(output_obj, scatter_gather_tensors=self.scatter_gather_tensors)
        output, label = pack_return_tensors(return_tensors) if len(return_tensors) > 0 else (None, None)
        if gpc.config.get("model") is not None and hasattr(gpc.config.model, "num_experts"):
            dist.all_reduce(accum_moe_loss, group=gpc.get_group(ParallelMode.PIPELINE))
            if accum_loss is not None:      # same convention as the training step: the returned loss contains the auxiliary term
                accum_loss = accum_loss + accum_moe_loss
            return output, label, accum_loss, accum_moe_loss
        return output, label, accum_loss

    def _forward_backward_step(self, engine, return_loss=True, return_output_label=True):
        M = self.num_microbatches
        pp_size, pp_rank = gpc.get_world_size(ParallelMode.PIPELINE), gpc.get_local_rank(ParallelMode.PIPELINE)
        num_warmup = min(pp_size - pp_rank - 1, M)
        num_1f1b = M - num_warmup
        input_objs, output_objs, moe_losses = [], [], []
        return_tensors = []
        accum_loss = torch.zeros(1, device=get_current_device()) if return_loss and gpc.is_pipeline_last_stage(True) else None
        accum_moe_loss = torch.zeros(1, device=get_current_device())
        ft = bt = self._tensor_shape
        sg = self.scatter_gather_tensors
        first, last = gpc.is_first_rank(ParallelMode.PIPELINE), gpc.is_last_rank(ParallelMode.PIPELINE)
        # ---- warm-up forwards
        for _ in range(num_warmup):
            if not first:
                if ft is None:
                    ft = comm.recv_obj_meta()
                input_obj = comm.recv_forward(ft, dtype=self.dtype, scatter_gather_tensors=sg)
            else:
                input_obj = None
            output_obj, moe_loss = self._forward_step(engine, input_obj, return_tensors, return_output_label,
                                                      accum_loss, accum_moe_loss)
            if not last:
                if self._tensor_shape is None:
                    bt = output_obj.shape
                    comm.send_obj_meta(output_obj)
                comm.send_forward(output_obj, scatter_gather_tensors=sg)
            input_objs.append(input_obj)
            output_objs.append(output_obj)
            moe_losses.append(moe_loss)
        if num_1f1b > 0:
            if not first:
                if ft is None:
                    ft = comm.recv_obj_meta()
                input_obj = comm.recv_forward(ft, dtype=self.dtype, scatter_gather_tensors=sg)
            else:
                input_obj = None
        # ---- steady state
        for i in range(num_1f1b):
            output_obj, moe_loss = self._forward_step(engine, input_obj, return_tensors, return_output_label,
                                                      accum_loss, accum_moe_loss)
            if last:
                output_obj_grad = None
            else:
                if self._tensor_shape is None:
                    bt = output_obj.shape
                    comm.send_obj_meta(output_obj)
                output_obj_grad = comm.send_forward_recv_backward(output_obj, bt, dtype=self.dtype,
                                                                  scatter_gather_tensors=sg)
            input_objs.append(input_obj)
            output_objs.append(output_obj)
            moe_losses.append(moe_loss)
            input_obj, output_obj, moe_loss = input_objs.pop(0), output_objs.pop(0), moe_losses.pop(0)
            input_obj_grad = self._backward_step(engine, i, input_obj, output_obj, output_obj_grad, moe_loss)
            if i == num_1f1b - 1:
                input_obj = None
                if not first:
                    comm.send_backward(input_obj_grad, scatter_gather_tensors=sg)
            else:
                if first:
                    input_obj = None
                else:
                    input_obj = comm.send_backward_recv_forward(input_obj_grad, ft, dtype=self.dtype,
                                                                scatter_gather_tensors=sg)
        # ---- cool-down backwards
        for i in range(num_warmup):
            input_obj, output_obj, moe_loss = input_objs.pop(0), output_objs.pop(0), moe_losses.pop(0)
            output_obj_grad = None if last else comm.recv_backward(bt, dtype=self.dtype, scatter_gather_tensors=sg)
            input_obj_grad = self._backward_step(engine, num_1f1b + i, input_obj, output_obj, output_obj_grad, moe_loss)
            if not first:
                comm.send_backward(input_obj_grad, scatter_gather_tensors=sg)
        output, label = pack_return_tensors(return_tensors) if len(return_tensors) > 0 else (None, None)
        if hasattr(gpc.config.model, "num_experts"):
            dist.all_reduce(accum_moe_loss, group=gpc.get_group(ParallelMode.PIPELINE))
            if accum_loss is not None:
                accum_loss = accum_loss + accum_moe_loss
            return output, label, accum_loss, accum_moe_loss
        return output, label, accum_loss

    @llm_timeout(func_name="nointerleaved_forward_backward_step")
    def forward_backward_step(self, engine, data_iter, forward_only=False, return_loss=True, return_output_label=True):
        assert forward_only or return_loss, "'return_loss' has to be True when 'forward_only' is False"
        if engine.optimizer is not None and hasattr(engine.optimizer, "wait_param_sync"):
            engine.optimizer.wait_param_sync()
        self.load_batch(engine, data_iter)
        if forward_only:
            with torch.no_grad():
                return self._forward_only_step(engine, return_loss, return_output_label)
        return self._forward_backward_step(engine, return_loss, return_output_label)


class InterleavedPipelineScheduler(PipelineScheduler):
    """Interleaved 1F1B over ``num_chunks`` virtual stages per rank (reference ``:711-1430``)."""

    def __init__(self, num_microbatches: int, num_chunks: int, dtype: torch.dtype = torch.float,
                 data_process_func: Callable = None, tensor_shape=None, scatter_gather_tensors: bool = False,
                 scheduler_hooks: Optional[List[SchedulerHook]] = None, communication_overlap: bool = False):
        assert num_microbatches % gpc.get_world_size(ParallelMode.PIPELINE) == 0, (
            "num_microbatches must be an integer multiple of pipeline parallel world size"
        )
        assert isinstance(num_chunks, int) and num_chunks > 0
        super().__init__(num_microbatches, dtype=dtype, data_process_func=data_process_func, tensor_shape=tensor_shape,
                         scatter_gather_tensors=scatter_gather_tensors, scheduler_hooks=scheduler_hooks)
        gpc.set_virtual_pipeline_parallel_size(num_chunks)
        gpc.set_virtual_pipeline_parallel_rank(0)
        self._num_chunks = num_chunks
        self._communication_overlap = communication_overlap
        self._pp_size = gpc.get_world_size(ParallelMode.PIPELINE)
        self._pp_rank = gpc.get_local_rank(ParallelMode.PIPELINE)
        self._clear_state()

    def _clear_state(self):
        n = self._num_chunks
        self._accum_loss = None
        self._accum_moe_loss = None
        self._return_tensors = None
        self._input_objs = [[] for _ in range(n)]
        self._output_objs = [[] for _ in range(n)]
        self._moe_losses = [[] for _ in range(n)]
        self._output_obj_grads = [[] for _ in range(n)]
        self.microbatch_offset = [0 for _ in range(n)]

    def load_batch(self, engine, data_iter):
        super().load_batch(engine, data_iter)
        self.microbatch_offset = [0 for _ in range(self._num_chunks)]

    def load_micro_batch(self, model_chunk_id):  # pylint: disable=W0221
        data, label = self._load_micro_batch(self.batch_data, self.batch_label, self.microbatch_offset[model_chunk_id],
                                             self.bsz_stride)
        self.microbatch_offset[model_chunk_id] += self.bsz_stride
        ms = data.pop("max_seqlen", None)
        data, label = move_to_device(data), move_to_device(label)
        if ms is not None:
            data["max_seqlen"] = ms
        return data, label

    def _chunk_of(self, k: int, forward: bool) -> int:
        """virtual-stage index of the k-th micro-step (reference ``:925-944``)."""
        idx = k % (self._pp_size * self._num_chunks) // self._pp_size
        return idx if forward else self._num_chunks - idx - 1

    def _chunk_model(self, engine, chunk_id):
        model = engine.model
        if isinstance(model, NaiveAMPModel):
            return lambda **kw: model.convert_to_fp32(model.model[chunk_id](**{k: model._convert_to_fp16(v) for k, v in kw.items()})) \
                if model._output_to_fp32 else model.model[chunk_id](**{k: model._convert_to_fp16(v) for k, v in kw.items()})
        return model[chunk_id]

    def _fwd(self, engine, chunk_id, input_obj=None):
        gpc.set_virtual_pipeline_parallel_rank(chunk_id)
        if gpc.is_pipeline_first_stage() and len(self._input_objs[chunk_id]) == len(self._output_objs[chunk_id]):
            self._input_objs[chunk_id].append(None)
        if input_obj is None:
            input_obj = self._input_objs[chunk_id][-1]
        data, label = self.load_micro_batch(chunk_id)
        data = self._prep_stage_input(input_obj, data)
        is_moe = hasattr(gpc.config.model, "num_experts")
        self._call_hooks("before_forward", data)
        out = self._chunk_model(engine, chunk_id)(**data)
        output_obj, moe_losses = out if is_moe else (out, None)
        self._call_hooks("after_forward", output_obj)
        moe_loss = None
        if gpc.is_pipeline_last_stage():
            self._call_hooks("post_helper_func", output_obj, label)
            if self._return_tensors is not None:
                self._return_tensors.append((output_obj, label))
            if self._accum_loss is not None:
                self._call_hooks("before_criterion", output_obj, label)
                loss = self._call_engine_criterion(engine, output_obj, label)
                self._call_hooks("after_criterion", loss)
                loss_reduced = loss / self.num_microbatches
                self._accum_loss.add_(loss_reduced.detach())
                output_obj = loss_reduced
        if is_moe and gpc.config.model.num_experts > 1 and moe_losses:
            moe_loss = sum(moe_losses) * gpc.config.loss.moe_loss_coeff / self.num_microbatches
            self._accum_moe_loss.add_(moe_loss.detach())
        self._output_objs[chunk_id].append(output_obj)
        self._moe_losses[chunk_id].append(moe_loss)
        return output_obj

    def _bwd(self, engine, chunk_id, step_id):
        gpc.set_virtual_pipeline_parallel_rank(chunk_id)
        if gpc.is_pipeline_last_stage() and len(self._output_obj_grads[chunk_id]) == 0:
            self._output_obj_grads[chunk_id].append(None)
        input_obj = self._input_objs[chunk_id].pop(0)
        output_obj = self._output_objs[chunk_id].pop(0)
        moe_loss = self._moe_losses[chunk_id].pop(0)
        output_obj_grad = self._output_obj_grads[chunk_id].pop(0)
        return self._backward_step(engine, step_id, input_obj, output_obj, output_obj_grad, moe_loss)

    def _run(self, engine, forward_only):
        """Megatron-style interleaved schedule. p2p exchanges are started right after the producing compute and
        collected right before the consuming compute, so with ``communication_overlap`` they hide behind it."""
        M, C, P, r = self.num_microbatches, self._num_chunks, self._pp_size, self._pp_rank
        total = M * C
        shape = self._tensor_shape
        assert shape is not None, "interleaved pipeline needs a static tensor_shape"
        sg = self.scatter_gather_tensors
        if forward_only:
            warm = total
        elif M == P:
            warm = total
        else:
            warm = min((P - r - 1) * 2 + (C - 1) * P, total)
        remaining = total - warm
        kw = dict(dtype=self.dtype, scatter_gather_tensors=sg)

        gpc.set_virtual_pipeline_parallel_rank(0)
        if not gpc.is_pipeline_first_stage():
            self._input_objs[0].append(comm.recv_forward(shape, **kw))
        # ---- warm-up
        for k in range(warm):
            chunk = self._chunk_of(k, True)
            out = self._fwd(engine, chunk)
            next_chunk = self._chunk_of(k + 1, True) if k + 1 < total else None
            recv_prev = next_chunk is not None
            if recv_prev:
                gpc.set_virtual_pipeline_parallel_rank(next_chunk)
                if gpc.is_pipeline_first_stage():
                    recv_prev = False
            gpc.set_virtual_pipeline_parallel_rank(chunk)
            send = None if gpc.is_pipeline_last_stage() else out
            if k == warm - 1 and not forward_only and remaining > 0:
                # last warm-up forward also posts the first backward receive
                bchunk = self._chunk_of(0, False)
                gpc.set_virtual_pipeline_parallel_rank(bchunk)
                recv_next = not gpc.is_pipeline_last_stage()
                inp, grad = comm.send_forward_backward_recv_forward_backward(
                    send, None, shape if recv_prev else None, shape if recv_next else None, **kw)
                if recv_next:
                    self._output_obj_grads[bchunk].append(grad)
            else:
                inp = comm.send_forward_recv_forward(send, shape if recv_prev else None, **kw)
            if recv_prev:
                self._input_objs[next_chunk].append(inp)
            if forward_only:
                # outputs are not needed for a backward pass
                self._input_objs[chunk].pop(0) if self._input_objs[chunk] else None
                self._output_objs[chunk].pop(0)
                self._moe_losses[chunk].pop(0)
        if forward_only:
            return
        # ---- steady 1F1B
        for k in range(remaining):
            fk = k + warm
            fchunk = self._chunk_of(fk, True)
            out = self._fwd(engine, fchunk)
            bchunk = self._chunk_of(k, False)
            in_grad = self._bwd(engine, bchunk, k)
            gpc.set_virtual_pipeline_parallel_rank(fchunk)
            send_f = None if gpc.is_pipeline_last_stage() else out
            gpc.set_virtual_pipeline_parallel_rank(bchunk)
            send_b = None if gpc.is_pipeline_first_stage() else in_grad
            next_f = self._chunk_of(fk + 1, True) if fk + 1 < total else None
            recv_prev = next_f is not None
            if recv_prev:
                gpc.set_virtual_pipeline_parallel_rank(next_f)
                recv_prev = not gpc.is_pipeline_first_stage()
            next_b = self._chunk_of(k + 1, False) if k + 1 < total else None
            recv_next = next_b is not None
            if recv_next:
                gpc.set_virtual_pipeline_parallel_rank(next_b)
                recv_next = not gpc.is_pipeline_last_stage()
            inp, grad = comm.send_forward_backward_recv_forward_backward(
                send_f, send_b, shape if recv_prev else None, shape if recv_next else None, **kw)
            if recv_prev:
                self._input_objs[next_f].append(inp)
            if recv_next:
                self._output_obj_grads[next_b].append(grad)
        # ---- cool-down
        if remaining == 0 and warm > 0:
            bchunk = self._chunk_of(0, False)
            gpc.set_virtual_pipeline_parallel_rank(bchunk)
            if not gpc.is_pipeline_last_stage():
                self._output_obj_grads[bchunk].append(comm.recv_backward(shape, **kw))
        for k in range(remaining, total):
            bchunk = self._chunk_of(k, False)
            in_grad = self._bwd(engine, bchunk, k)
            gpc.set_virtual_pipeline_parallel_rank(bchunk)
            send_b = None if gpc.is_pipeline_first_stage() else in_grad
            next_b = self._chunk_of(k + 1, False) if k + 1 < total else None
            recv_next = next_b is not None
            if recv_next:
                gpc.set_virtual_pipeline_parallel_rank(next_b)
                recv_next = not gpc.is_pipeline_last_stage()
            grad = comm.send_backward_recv_backward(send_b, shape if recv_next else None, **kw)
            if recv_next:
                self._output_obj_grads[next_b].append(grad)

    def _backward_step(self, engine, step_id, input_obj, output_obj, output_obj_grad, moe_loss=None):
        if engine.optimizer is not None:
            engine.optimizer.skip_grad_reduce = step_id != self.num_microbatches * self._num_chunks - 1
        if input_obj is not None and isinstance(input_obj, torch.Tensor):
            input_obj.retain_grad()
        self._call_hooks("before_backward", output_obj, output_obj_grad)
        if moe_loss is None or not moe_loss.requires_grad:
            if output_obj_grad is None:
                engine.backward(output_obj)
            else:
                engine.backward_by_grad(output_obj, output_obj_grad)
        else:
            scale = engine.optimizer.grad_scaler.scale if hasattr(engine.optimizer, "grad_scaler") else 1.0
            if output_obj_grad is None:
                engine.backward(output_obj + moe_loss)
            else:
                engine.backward_by_grad([output_obj, moe_loss * scale], [output_obj_grad, None])
        self._call_hooks("after_backward", None)
        return input_obj.grad if isinstance(input_obj, torch.Tensor) else None

    @llm_timeout(func_name="interleaved_forward_backward_step")
    def forward_backward_step(self, engine, data_iter, forward_only=False, return_loss=True, return_output_label=True):
        assert forward_only or return_loss
        if engine.optimizer is not None and hasattr(engine.optimizer, "wait_param_sync"):
            engine.optimizer.wait_param_sync()
        gpc.set_virtual_pipeline_parallel_rank(0)
        self.load_batch(engine, data_iter)
        self._clear_state()
        if return_loss and gpc.is_pipeline_last_stage(ignore_virtual=True):
            self._accum_loss = torch.zeros(1, device=get_current_device())
        self._accum_moe_loss = torch.zeros(1, device=get_current_device())
        if return_output_label:
            self._return_tensors = []
        if forward_only:
            with torch.no_grad():
                self._run(engine, True)
        else:
            self._run(engine, False)
        output, label = pack_return_tensors(self._return_tensors) if self._return_tensors else (None, None)
        accum_loss, accum_moe = self._accum_loss, self._accum_moe_loss
        is_moe = hasattr(gpc.config.model, "num_experts")
        if is_moe:
            dist.all_reduce(accum_moe, group=gpc.get_group(ParallelMode.PIPELINE))
            if accum_loss is not None:
                accum_loss = accum_loss + accum_moe
        self._clear_state()
        gpc.set_virtual_pipeline_parallel_rank(0)
        if is_moe:
            return output, label, accum_loss, accum_moe
        return output, label, accum_loss
