#!/usr/bin/env python
"""Config-driven training entry point (same CLI and flow as the reference ``train.py``):

    torchrun --nproc_per_node=8 --master-addr 127.0.0.1 train.py --config configs/7B_internlm2.py --launcher torch
    srun -n64 python train.py --config configs/7B_internlm2.py --launcher slurm
"""
import socket
import time
import traceback
from functools import partial

import torch.distributed as dist

import internevo_b200
from internevo_b200.checkpoint import CheckpointManager
from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.data import build_train_loader_with_data_type, build_valid_loader_with_data_type
from internevo_b200.data.train_state import get_train_state
from internevo_b200.eval.evaluation import evaluate_on_val_dls
from internevo_b200.initialize import initialize_distributed_env
from internevo_b200.models.losses import FlashGPTLMLoss
from internevo_b200.models.metrics import AccPerplex
from internevo_b200.monitor import initialize_monitor_manager, send_alert_message
from internevo_b200.monitor.monitor import monitor_manager as mm
from internevo_b200.train import (
    get_scheduler_hooks,
    initialize_isp_communicator,
    initialize_llm_profile,
    initialize_model,
    initialize_optimizer,
    load_new_batch,
    record_current_batch_training_metrics,
)
from internevo_b200.utils.common import (
    BatchSkipper,
    enable_pytorch_expandable_segments,
    get_current_device,
    get_megatron_flops,
    launch_time,
    parse_args,
)
from internevo_b200.utils.gputest import empty_cache_and_diag
from internevo_b200.utils.logger import get_logger, initialize_uniscale_logger
from internevo_b200.utils.megatron_timers import megatron_timer as timer
from internevo_b200.utils.parallel import get_parallel_log_file_name
from internevo_b200.utils.simple_memory_profiler import SimpleMemoryProfiler
from internevo_b200.utils.writer import Writer

logger = get_logger(__file__)


def main(args):
    enable_pytorch_expandable_segments()
    skip_batches = gpc.config.data.skip_batches
    total_steps = gpc.config.data.total_steps
    valid_every = gpc.config.data.valid_every
    label_smoothing = gpc.config.loss.label_smoothing

    get_tflops_func = partial(
        get_megatron_flops,
        checkpoint=gpc.config.model.checkpoint,
        seq_len=gpc.config.data["seq_len"],
        hidden_size=gpc.config.model.hidden_size,
        num_layers=gpc.config.model.num_layers,
        vocab_size=gpc.config.model.vocab_size,
        global_batch_size=gpc.config.data.micro_bsz * gpc.config.data.micro_num * gpc.get_world_size(ParallelMode.DATA),
        global_world_size=gpc.get_world_size(ParallelMode.GLOBAL),
        mlp_ratio=gpc.config.model["mlp_ratio"],
    )

    # one launch-time string for the whole job (log / tensorboard folder names)
    current_time = launch_time()
    if gpc.is_distributed and gpc.get_world_size(ParallelMode.GLOBAL) > 1:
        objs = [current_time]
        dist.broadcast_object_list(objs, src=0)
        current_time = objs[0]
    current_time = current_time.replace(":", ".")
    initialize_uniscale_logger(job_name=gpc.config.JOB_NAME, launch_time=current_time,
                               file_name=get_parallel_log_file_name()) if gpc.config.get("enable_file_log", False) else None

    model = initialize_model()
    isp_communicator = initialize_isp_communicator(model)
    try:
        with open(args.config, "r") as f:
            config_lines = f.readlines()
    except (OSError, TypeError):
        config_lines = []

    criterion = FlashGPTLMLoss(parallel_output=gpc.config.model.get("parallel_output", True), label_smoothing=label_smoothing)
    train_dl, dataset_types = build_train_loader_with_data_type()
    val_dls = build_valid_loader_with_data_type() if valid_every > 0 else None
    train_state = get_train_state(train_dl)
    optimizer, beta2_scheduler, lr_scheduler = initialize_optimizer(model, isp_communicator)

    ckpt_manager = CheckpointManager(
        ckpt_config=gpc.config.ckpt, model=model, optimizer=optimizer, lr_scheduler=lr_scheduler, train_dl=train_dl,
        model_config=gpc.config.model, model_config_file="".join(config_lines),
        feishu_address=gpc.config.monitor.alert.get("feishu_alert_address", None),
    )
    ckpt_manager.try_resume_training(train_state, current_time)

    writer = Writer(
        job_name=gpc.config.JOB_NAME, launch_time=current_time, file_name=get_parallel_log_file_name(),
        tensorboard_folder=gpc.config.tensorboard_folder, resume_tb_folder=train_state.resume_tb_folder,
        step_count=train_state.step_count, config=None, logger=logger, enable_tb=gpc.config.enable_tb,
        queue_max_length=gpc.config.tensorboard.queue_max_length, total_steps=total_steps,
    )
    metric = AccPerplex(device=get_current_device(), tp_pg=gpc.get_group(ParallelMode.TENSOR),
                        dp_pg=gpc.get_group(ParallelMode.DATA), dataset_types=dataset_types)

    trainer, train_dl, _, _ = internevo_b200.initialize_trainer(
        model=model, optimizer=optimizer, criterion=criterion, train_dataloader=train_dl, lr_scheduler=lr_scheduler,
        beta2_scheduler=beta2_scheduler, scheduler_hooks=get_scheduler_hooks(metric, optimizer, isp_communicator),
    )

    memory_profiler = None
    if args.profiling:
        memory_profiler = SimpleMemoryProfiler(
            model, optimizer,
            log_folder=f"RUN/{gpc.config.JOB_NAME}/{current_time}/memory_trace/rank{gpc.get_global_rank()}_"
            f"dp{gpc.get_local_rank(ParallelMode.DATA)}_wp{gpc.get_local_rank(ParallelMode.WEIGHT)}_"
            f"tp{gpc.get_local_rank(ParallelMode.TENSOR)}",
        )
    batch_skipper = BatchSkipper(skip_batches)
    trainer.train()
    train_iter = iter(train_dl)

    with initialize_llm_profile(profiling=args.profiling, start_time=current_time) as prof:
        for batch_count in range(train_state.batch_count, total_steps):
            empty_cache_and_diag(batch_count, interval=gpc.config.data.empty_cache_and_diag_interval)
            start_time = time.time()
            timer("one-batch").start()
            batch, train_iter = load_new_batch(train_dl=train_dl, train_iter=train_iter, train_state=train_state)
            train_state.batch_count = batch_count
            train_state.num_consumed_samples_in_epoch += len(batch[1])
            if batch_skipper(batch_count):
                if gpc.is_rank_for_log():
                    logger.info(f"Skip batch count:`{batch_count}`...")
                timer("one-batch").stop()
                continue

            # dataset type of every token: feeds the per-dataset loss / accuracy metrics, not the model
            type_ids = batch[0].pop("type_ids", None)
            if type_ids is not None and metric is not None:
                metric.set_current_type_ids(type_ids=type_ids)

            trainer.zero_grad()
            timer("fwd-bwd").start()
            moe_loss = None
            if hasattr(gpc.config.model, "num_experts"):
                _, _, loss, moe_loss = trainer.execute_schedule(batch, forward_only=False, return_loss=True,
                                                                return_output_label=False)
            else:
                _, _, loss = trainer.execute_schedule(batch, forward_only=False, return_loss=True,
                                                      return_output_label=False)
            timer("fwd-bwd").stop()

            success_update, grad_norm_groups = trainer.step()
            if success_update:
                train_state.step_count += 1
            else:
                train_state.inf_nan_skip_batches += 1
                if -1 in grad_norm_groups.values() and gpc.is_rank_for_log():
                    logger.warning(f"Warning: skip parameter update at step {batch_count}.")
                    send_alert_message(address=gpc.config.monitor.alert.get("feishu_alert_address", None),
                                       message=f"Warning: skip parameter update at step {batch_count}.")

            record_current_batch_training_metrics(
                get_tflops_func=get_tflops_func, logger=logger, writer=writer, success_update=success_update,
                batch_count=batch_count, batch=batch, train_state=train_state, optimizer=optimizer,
                beta2_scheduler=beta2_scheduler, trainer=trainer, start_time=start_time, loss=loss, moe_loss=moe_loss,
                grad_norm=grad_norm_groups, metric=metric, update_panel=False,
            )
            timer("one-batch").stop()

            if valid_every > 0 and train_state.step_count % valid_every == 0 and val_dls is not None:
                evaluate_on_val_dls(trainer=trainer, val_dls=val_dls, writer=writer, logger=logger,
                                    step_count=train_state.step_count)

            if ckpt_manager.try_save_checkpoint(train_state):
                break
            if memory_profiler is not None:
                memory_profiler.step()
            if batch_count % 2 == 0:
                prof.step()

    ckpt_manager.wait_async_upload_finish()
    writer.close()


if __name__ == "__main__":
    args = parse_args()
    hostname = socket.gethostname()
    initialize_distributed_env(config=args.config, launcher=args.launcher, master_port=args.port, seed=args.seed,
                               backend=args.backend)
    assert hasattr(gpc, "config") and gpc.config is not None
    alert = gpc.config.monitor.alert
    with initialize_monitor_manager(job_name=gpc.config.JOB_NAME,
                                    alert_address=alert.get("feishu_alert_address") if alert.get("enable_feishu_alert") else None):
        failed = False
        try:
            main(args)
        except Exception:
            failed = True
            logger.error(f"Raise exception from {hostname} with rank id: {gpc.get_global_rank()}\n{traceback.format_exc()}")
            mm.monitor_exception(alert_address=alert.get("feishu_alert_address", None), excp_info=traceback.format_exc())
            raise
        finally:
            # after an exception the peers may be blocked in a collective: leave without a barrier so the job fails fast
            gpc.destroy(graceful=not failed)
