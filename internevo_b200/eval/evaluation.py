"""Validation loop (reference ``internlm/eval/evaluation.py:18-143``): forward-only schedule over every validation
loader with accuracy / perplexity / per-dataset loss, pipeline-aware."""
from contextlib import contextmanager

import torch
from tqdm import tqdm

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.models.metrics import AccPerplex, SchedulerMetricHook


@contextmanager
def switch_evaluation_no_pipeline_scheduler(trainer, grad_accum_size, metric_hook_list):
    if not gpc.is_using_parallel_mode(ParallelMode.PIPELINE):
        prev_data_process_func = trainer.schedule.data_process_func
        prev_grad_accum_size = trainer.schedule._grad_accum_size
        prev_metric_hooks = trainer.schedule._hooks
        try:
            trainer.schedule.data_process_func = None     # validation batches are not packed: nothing to un-pack
            trainer.schedule._grad_accum_size = grad_accum_size
            trainer.schedule._hooks = metric_hook_list
            yield
        finally:
            trainer.schedule.data_process_func = prev_data_process_func
            trainer.schedule._grad_accum_size = prev_grad_accum_size
            trainer.schedule._hooks = prev_metric_hooks
    else:
        yield


@contextmanager
def switch_evaluation_pipeline_scheduler(trainer, num_microbatches, tensor_shape, metric_hook_list):
    if gpc.is_using_parallel_mode(ParallelMode.PIPELINE):
        pre_data_process_func = trainer.schedule.data_process_func
        prev_num_microbatches = trainer.schedule.num_microbatches
        prev_tensor_shape = trainer.schedule.tensor_shape
        prev_metric_hooks = trainer.schedule._hooks
        try:
            trainer.schedule.data_process_func = None
            trainer.schedule.num_microbatches = num_microbatches
            trainer.schedule.tensor_shape = tensor_shape
            trainer.schedule._hooks = metric_hook_list
            yield
        finally:
            trainer.schedule.data_process_func = pre_data_process_func
            trainer.schedule.num_microbatches = prev_num_microbatches
            trainer.schedule.tensor_shape = prev_tensor_shape
            trainer.schedule._hooks = prev_metric_hooks
    else:
        yield


@contextmanager
def switch_evaluation_mode(trainer, metric_hook_list=None):
    """``gpc.is_evaluating`` for the duration of the block.  Without ``metric_hook_list`` (this framework's own validation
    loop) the trainer is also put into eval mode and back; with it the call has the reference's semantics
    (``eval/evaluation.py:28-42``): the schedule's hooks are replaced by the list and its data_process_func is disabled."""
    prev = gpc.is_evaluating
    if metric_hook_list is None:
        try:
            gpc.is_evaluating = True
            trainer.eval()
            yield
        finally:
            gpc.is_evaluating = prev
            trainer.train()
        return
    prev_func, prev_hooks = trainer.schedule.data_process_func, trainer.schedule._hooks
    try:
        gpc.is_evaluating = True
        trainer.schedule.data_process_func = None
        trainer.schedule._hooks = metric_hook_list
        yield
    finally:
        gpc.is_evaluating = prev
        trainer.schedule.data_process_func = prev_func
        trainer.schedule._hooks = prev_hooks


def evaluate_on_val_dls(trainer, val_dls, writer, logger, step_count, update_panel: bool = False, streaming: bool = False):
    with switch_evaluation_mode(trainer):
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        criterion = trainer.engine.criterion
        data_cfg = gpc.config.data
        for val_name, val_dl in val_dls.items():
            if not streaming and len(val_dl) == 0 and gpc.is_rank_for_log():
                logger.info(f"Validation dataset: {val_name} is empty")
                continue
            val_metric = AccPerplex(dataset_types=None)
            hook = SchedulerMetricHook(metric=val_metric, skip=True, criterion=criterion)
            val_loss, n = 0.0, 0
            it = enumerate(val_dl)
            if gpc.is_rank_for_log():
                it = tqdm(it, desc="Val.", total=len(val_dl) if not streaming else None, position=1, leave=False)
            for _, batch in it:
                with torch.inference_mode():
                    total_bsz = len(batch[1])
                    assert total_bsz % data_cfg.micro_bsz == 0
                    num_micro, rows = total_bsz // data_cfg.micro_bsz, data_cfg.micro_bsz
                    pp = gpc.get_world_size(ParallelMode.PIPELINE)
                    if pp > 1 and getattr(trainer.schedule, "_num_chunks", 1) > 1 and num_micro % pp != 0:
                        # the interleaved schedule runs micro-batches in groups of pp: cut the validation batch into the
                        # smallest multiple of pp micro-batches that divides it (fewer rows per micro-batch)
                        fits = [m for m in range(pp, total_bsz + 1, pp) if total_bsz % m == 0]
                        assert fits, (f"interleaved pipeline (pp = {pp}) cannot split a validation batch of {total_bsz} samples: "
                                      f"make valid_micro_num * micro_bsz a multiple of {pp}")
                        num_micro, rows = fits[0], total_bsz // fits[0]
                    sp = gpc.get_world_size(ParallelMode.TENSOR) if gpc.config.parallel.get("sequence_parallel", False) else 1
                    shape = (rows * batch[0]["input_ids"].shape[1] // sp, gpc.config.model["hidden_size"])
                    if gpc.is_using_parallel_mode(ParallelMode.PIPELINE):
                        with switch_evaluation_pipeline_scheduler(trainer, num_micro, shape, [hook]):
                            trainer.schedule.bsz_stride = rows
                            out = trainer.execute_schedule(batch, forward_only=True, return_loss=True, return_output_label=False)
                    else:
                        with switch_evaluation_no_pipeline_scheduler(trainer, num_micro, [hook]):
                            out = trainer.execute_schedule(batch, forward_only=True, return_loss=True, return_output_label=False)
                    loss = out[2]
                    moe_loss = out[3] if len(out) > 3 else None
                if gpc.is_no_pp_or_last_stage() and loss is not None:
                    # language-model loss only: the schedulers add the MoE auxiliary loss to what they return (reference ``:107``)
                    val_loss += float(loss) - (float(moe_loss) if moe_loss is not None else 0.0)
                    n += 1
            if n > 0:
                res = val_metric.get_metric()
                val_loss = val_loss / (n + 1e-6)      # the reference's divisor (``eval/evaluation.py:115``)
                if gpc.is_rank_for_log():
                    infos = {"step": step_count, f"val/{val_name}_loss": val_loss, f"val/{val_name}_acc": res["acc"],
                             f"val/{val_name}_plex": res["perplexity"]}
                    for k, v in infos.items():
                        if k != "step" and writer is not None:
                            writer.add_scalar(key=k, value=v, step=step_count)
                    logger.info("Validation on {}: ".format(val_name) + " ".join(f"{k}={v}" for k, v in infos.items()))
        trainer.train()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        if gpc.is_distributed:
            torch.distributed.barrier()
