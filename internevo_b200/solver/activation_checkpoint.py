"""Activation checkpointing that restores every per-``ParallelMode`` RNG stream for the recompute (reference
``internlm/solver/activation_checkpoint.py:40-152``), with optional CPU offload of the saved inputs."""
from __future__ import annotations


import torch
from torch.utils.checkpoint import check_backward_validity, detach_variable

from internevo_b200.core.context import global_context as gpc  # noqa: F401
from internevo_b200.core.context.random import get_current_mode, get_states, set_mode, set_seed_states, sync_states


def _dev_rng_get():
    return torch.cuda.get_rng_state() if torch.cuda.is_available() else torch.get_rng_state()


def _dev_rng_set(s):
    if torch.cuda.is_available():
        torch.cuda.set_rng_state(s)
    else:
        torch.set_rng_state(s)


def copy_to_device(obj, device):
    if torch.is_tensor(obj):
        ret = obj.to(device).detach()
        ret.requires_grad = obj.requires_grad
        return ret
    if isinstance(obj, list):
        return [copy_to_device(i, device) for i in obj]
    if isinstance(obj, tuple):
        return tuple(copy_to_device(v, device) for v in obj)
    if isinstance(obj, dict):
        return {k: copy_to_device(v, device) for k, v in obj.items()}
    return obj


class CheckpointFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, run_function, activation_offload=False, *args):
        check_backward_validity(args)
        ctx.run_function = run_function
        ctx.activation_offload = activation_offload
        ctx.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        ctx.fwd_cpu_rng_state = torch.get_rng_state()
        sync_states()
        ctx.fwd_seed_states = get_states(copy=True)
        ctx.fwd_current_mode = get_current_mode()
        ctx.had_autocast_in_fwd = torch.is_autocast_enabled() if hasattr(torch, "is_autocast_enabled") else False
        inputs_cuda = copy_to_device(args, ctx.device) if activation_offload else args
        with torch.no_grad():
            outputs = run_function(*inputs_cuda)
        ctx.inputs, ctx.tensor_indices = [], []
        tensor_inputs = []
        for i, arg in enumerate(args):
            if torch.is_tensor(arg):
                tensor_inputs.append(copy_to_device(arg, "cpu") if activation_offload else arg)
                ctx.tensor_indices.append(i)
                ctx.inputs.append(None)
            else:
                ctx.inputs.append(arg)
        if activation_offload:
            ctx.tensor_inputs = tensor_inputs
        else:
            ctx.save_for_backward(*tensor_inputs)
        return outputs

    @staticmethod
    def backward(ctx, *args):
        if not torch.autograd._is_checkpoint_valid():
            raise RuntimeError("Checkpointing is not compatible with .grad() or when an `inputs` parameter is passed")
        inputs = list(ctx.inputs)
        tensors = ctx.tensor_inputs if ctx.activation_offload else ctx.saved_tensors
        bwd_cpu_rng_state = torch.get_rng_state()
        sync_states()
        bwd_seed_states = get_states(copy=True)
        bwd_current_mode = get_current_mode()
        torch.set_rng_state(ctx.fwd_cpu_rng_state)
        for mode, state in ctx.fwd_seed_states.items():
            set_seed_states(mode, state)
        if ctx.fwd_current_mode is not None:
            set_mode(ctx.fwd_current_mode)
        if ctx.activation_offload:
            tensors = copy_to_device(tensors, ctx.device)
        for i, idx in enumerate(ctx.tensor_indices):
            inputs[idx] = tensors[i]
        detached_inputs = detach_variable(tuple(inputs))
        with torch.enable_grad():
            outputs = ctx.run_function(*detached_inputs)
        if isinstance(outputs, torch.Tensor):
            outputs = (outputs,)
        torch.set_rng_state(bwd_cpu_rng_state)
        for mode, state in bwd_seed_states.items():
            set_seed_states(mode, state)
        if bwd_current_mode is not None:
            set_mode(bwd_current_mode)
        outs, grads = [], []
        for i, o in enumerate(outputs):
            if torch.is_tensor(o) and o.requires_grad and args[i] is not None:
                outs.append(o)
                grads.append(args[i])
        if not outs:
            raise RuntimeError("none of output has requires_grad=True, this checkpoint() is not necessary")
        torch.autograd.backward(outs, grads)
        g = tuple(inp.grad if isinstance(inp, torch.Tensor) else None for inp in detached_inputs)
        return (None, None) + g


def activation_checkpoint(function, activation_offload, *args, use_reentrant: bool = True):
    """Checkpoint ``function(*args)``; reentrant mode uses ``CheckpointFunction`` above, otherwise saved-tensor hooks."""
    if use_reentrant:
        return CheckpointFunction.apply(function, activation_offload, *args)
    return torch.utils.checkpoint.checkpoint(function, *args, use_reentrant=False)
