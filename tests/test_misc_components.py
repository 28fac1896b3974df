"""Components that had no dedicated test: shared-module gradient handler (C12), logger (F2), gputest helpers (F5),
SimpleMemoryProfiler (F6), init functions (G3), web demo page (T3)."""
import logging
import math
import os
import sys

import torch

from common import run_distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------------------------- C12
def _shared_module_grads(rank, world):
    import torch.distributed as dist

    from internevo_b200.core.gradient_handler import PipelineSharedModuleGradientHandler

    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = torch.nn.Linear(4, 3)
    model.weight.pipeline_shared_module_pg = dist.group.WORLD      # tied across "stages"
    model.weight.grad = torch.full_like(model.weight, float(rank + 1))
    model.bias.grad = torch.full_like(model.bias, float(rank + 1))  # not tagged: untouched
    PipelineSharedModuleGradientHandler(model, None).handle_gradient()
    out = (model.weight.grad.clone(), model.bias.grad.clone())
    dist.destroy_process_group()
    return out


def test_pipeline_shared_module_gradient_handler_sums_tagged_grads():
    for rank, (gw, gb) in enumerate(run_distributed(_shared_module_grads, 2)):
        assert torch.all(gw == 3.0)              # 1 + 2 summed over the shared group
        assert torch.all(gb == float(rank + 1))


# ----------------------------------------------------------------------------------------------------------------- F2
def test_logger_singleton_file_sink_and_extra_handler(tmp_path):
    from internevo_b200.utils import logger as L

    lg = L.get_logger("whatever")
    assert lg is L.get_logger() and lg.name == L.LOGGER_NAME and not lg.propagate
    n_before = len(lg.handlers)
    assert n_before >= 1 and any(type(h) is logging.StreamHandler for h in lg.handlers)
    L.get_logger()
    L.get_logger("x")
    assert len(lg.handlers) == n_before           # repeated get_logger calls never stack handlers

    seen = []

    class Sink(logging.Handler):
        def emit(self, record):
            seen.append(record.getMessage())

    sink = Sink()
    L.add_handler(sink)
    lg2 = L.initialize_uniscale_logger(job_name="job", launch_time="t0", file_name="rank0", file_path=str(tmp_path))
    lg2.info("hello sinks")
    for h in lg2.handlers:
        h.flush()
    assert "hello sinks" in seen
    assert "hello sinks" in open(tmp_path / "rank0.log").read()
    lg.removeHandler(sink)
    for h in list(lg.handlers):
        if isinstance(h, logging.FileHandler):
            lg.removeHandler(h)
            h.close()


# ----------------------------------------------------------------------------------------------------------------- F5
def test_gputest_helpers_cpu():
    from internevo_b200.utils import gputest

    # attention TFLOPS formula used by the micro-benchmark: 4 * b * s^2 * h * d / time / 1e12
    assert math.isclose(gputest.flops(2, 128, 64, 8, 1.0), 4 * 2 * 128 ** 2 * 8 * 64 / 1e12, rel_tol=1e-9)
    assert isinstance(gputest.get_cpu_temperature(), (int, float))
    assert isinstance(gputest.get_gpu_temperature(), (int, float))   # -1 when no NVML device is present


# ----------------------------------------------------------------------------------------------------------------- F6
def test_simple_memory_profiler_accounts_params_grads_optimizer_and_activations(tmp_path):
    from internevo_b200.utils.simple_memory_profiler import SimpleMemoryProfiler, SimpleMemState

    st = SimpleMemState("root")
    st.add("a.b", 10)
    st.add("a.c", 6)
    st.add("d", 4)
    assert st.total_mem == 20 and "a" in st.dump() and st.to_json()["name"] == "root"

    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))

    class G:
        name = "default"
        master = torch.zeros(10)
        exp_avg = torch.zeros(10)
        exp_avg_sq = torch.zeros(10)

    class Opt:
        groups = [G()]

    prof = SimpleMemoryProfiler(model, Opt(), str(tmp_path), total_steps=2)
    n_param_bytes = sum(p.numel() * p.element_size() for p in model.parameters())
    assert prof._param.total_mem == n_param_bytes == prof._grad.total_mem
    assert prof._os.total_mem == 3 * 10 * 4
    for _ in range(2):
        model(torch.randn(5, 8)).sum().backward()
        prof.step()
    text = open(tmp_path / "memory_2.log").read()
    assert "parameters" in text and "activations" in text and "optimizer_states" in text
    assert prof._stoped and prof._activation.total_mem == (5 * 16 + 5 * 16 + 5 * 4) * 4
    model(torch.randn(5, 8))                      # hooks are removed after the last profiled step
    assert prof._activation.total_mem == (5 * 16 + 5 * 16 + 5 * 4) * 4
    # everything the forward kept alive was handed back during backward; the peak saw all three outputs at once
    assert prof._activation.live == 0 and prof._activation.peak_live == (5 * 16 + 5 * 16 + 5 * 4) * 4
    assert "alive at the peak" in text and os.path.exists(tmp_path / "memory.html")


def test_memory_profiler_tracks_model_chunks_and_in_place_consumers(tmp_path):
    """Interleaved pipeline stages: one activation tree per chunk; an in-place op on a profiled module's output (the attention
    block rotates the wqkv projection in place) must keep working while the profiler is attached."""
    from internevo_b200.utils.simple_memory_profiler import ActivationMemState, SimpleMemoryProfiler

    class Wrapped(torch.nn.Module):          # stands for NaiveAMPModel: the chunk is ``.model``
        def __init__(self, m):
            super().__init__()
            self.model = m

        def forward(self, x):
            return self.model(x)

    chunks = torch.nn.ModuleList([Wrapped(torch.nn.Linear(4, 4)), Wrapped(torch.nn.Linear(4, 2))])
    prof = SimpleMemoryProfiler(chunks, None, str(tmp_path), total_steps=1)
    assert isinstance(prof._activation, ActivationMemState) and len(prof._activation.states) == 2
    h = chunks[0](torch.randn(3, 4))
    h.mul_(2.0)                              # in place on the hooked output
    chunks[1](h).sum().backward()
    assert prof._activation.inited == [True, True]
    assert [s.total_mem for s in prof._activation.states] == [3 * 4 * 4, 3 * 2 * 4]
    assert prof._activation.live == 0
    prof.step()
    text = open(tmp_path / "memory_1.log").read()
    assert "activations_0" in text and "activations_1" in text and "chunk1" in text


# ----------------------------------------------------------------------------------------------------------------- G3
def test_init_functions_statistics():
    from internevo_b200.initialize.initialize_tensor import (normal_, scaled_init_method_normal,
                                                              scaled_init_method_uniform, uniform_)

    torch.manual_seed(0)
    t = torch.empty(400, 400)
    normal_(std=0.02)(t)
    assert abs(float(t.std()) - 0.02) < 1e-3 and abs(float(t.mean())) < 1e-3
    scaled_init_method_normal(sigma=0.02, num_layers=8)(t)
    assert abs(float(t.std()) - 0.02 / 4.0) < 5e-4           # sigma / sqrt(2 * num_layers)
    uniform_(std=0.03)(t)
    a = math.sqrt(3 * 0.03)
    assert float(t.max()) <= a and float(t.min()) >= -a and float(t.max()) > 0.95 * a
    scaled_init_method_uniform(sigma=0.02, num_layers=2)(t)
    a = math.sqrt(3 * 0.02 / 2.0)
    assert float(t.abs().max()) <= a and float(t.abs().max()) > 0.95 * a


# ----------------------------------------------------------------------------------------------------------------- T3
def test_web_demo_serves_chat_page():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import web_demo
        from fastapi.testclient import TestClient
    except Exception as e:  # httpx (TestClient) may be missing in the image: check the route table instead
        import web_demo

        paths = {r.path for r in web_demo.app.routes}
        assert "/" in paths and "/v1/chat/completions" in paths, (paths, e)
        assert "/v1/chat/completions" in web_demo.PAGE
        return
    r = TestClient(web_demo.app).get("/")
    assert r.status_code == 200 and "internevo_b200 chat" in r.text


def test_nvtx_ranges_are_noops_without_cuda_and_toggle():
    from internevo_b200.utils import nvtx

    nvtx.enable(True)
    with nvtx.nvtx_range("cpu-safe"):       # no CUDA here: must not raise
        pass
    assert nvtx.enabled() == torch.cuda.is_available()
    nvtx.enable(False)
    assert not nvtx.enabled()


def test_accelerator_facade_cpu_answers_and_alias():
    """`internlm.accelerator.get_accelerator()` call surface (reference internlm/accelerator): one implementation, CPU-safe."""
    import internlm  # noqa: F401  (alias package)
    from internlm.accelerator import AcceleratorType, get_accelerator

    acc = get_accelerator()
    assert acc is get_accelerator()
    if not torch.cuda.is_available():
        assert acc.get_accelerator_backend() == AcceleratorType.CPU and acc.get_backend_name() == "cpu"
        assert acc.device_name() == "cpu" and acc.device_count() == 0 and acc.memory_allocated() == 0
        assert acc.communication_backend_name() == "gloo"
    else:
        assert acc.get_accelerator_backend() == AcceleratorType.GPU and acc.communication_backend_name() == "nccl"
    acc.synchronize()
    acc.empty_cache()
    acc.manual_seed(3)
    st = acc.get_rng_state()
    acc.set_rng_state(st)
    assert acc.is_bf16_supported() and acc.FloatTensor([1.0, 2.0]).dtype == torch.float32


def test_public_api_surface_importable_through_internlm_alias():
    """the names user scripts import from `internlm.*` resolve to this framework."""
    import importlib

    surface = {
        "internlm": ["initialize_trainer", "launch_from_torch", "launch_from_slurm", "get_default_parser"],
        "internlm.initialize": ["initialize_distributed_env", "try_bind_numa"],
        "internlm.core.context": ["global_context", "ParallelMode", "Config", "IS_TENSOR_ZERO_PARALLEL",
                                  "IS_REPLICA_ZERO_PARALLEL", "IS_WEIGHT_ZERO_PARALLEL", "IS_TENSOR_DATA_PARALLEL",
                                  "IS_TENSOR_EXPERT_DATA_PARALLEL", "set_mode", "get_seeds", "get_states", "seed",
                                  "sync_states", "add_seed", "get_current_mode"],
        "internlm.core.communication": ["recv_forward", "recv_backward", "send_forward", "send_backward",
                                        "send_forward_recv_backward", "send_backward_recv_forward", "send_obj_meta",
                                        "recv_obj_meta"],
        "internlm.core.scheduler": ["BaseScheduler", "NonPipelineScheduler", "PipelineScheduler",
                                    "InterleavedPipelineScheduler"],
        "internlm.train": ["initialize_model", "initialize_optimizer", "initialize_isp_communicator", "get_scheduler_hooks",
                           "load_new_batch", "record_current_batch_training_metrics", "initialize_llm_profile",
                           "set_fp32_attr_for_model", "set_parallel_attr_for_param_groups", "wrap_FSDP_model"],
        "internlm.model": ["MHA", "FeedForward", "Embedding1D", "RotaryEmbedding", "MoE", "ScaleColumnParallelLinear",
                           "BaseScaleColumnParallelLinear", "RewardModelLinear", "AccPerplex", "build_model_with_cfg",
                           "build_model_with_moe_cfg", "gather_forward_split_backward"],
        "internlm.solver": ["HybridZeroOptimizer", "Beta2Scheduler", "FineTuneCosineAnnealingWarmupLR"],
        "internlm.checkpoint": ["CheckpointManager"],
        "internlm.data": ["build_train_loader_with_data_type", "build_valid_loader_with_data_type"],
        "internlm.monitor": ["initialize_monitor_manager", "send_alert_message", "send_heartbeat", "set_env_var",
                             "initialize_light_monitor"],
        "internlm.utils.registry": ["MODEL_INITIALIZER"],
        "internlm.utils.common": ["SchedulerHook", "BatchSkipper", "get_megatron_flops", "parse_args", "launch_time",
                                  "get_current_device", "DummyProfile"],
        "internlm.accelerator": ["get_accelerator", "AcceleratorType"],
        "internlm.apis.inference": ["SequenceGenerator"],
        "internlm.utils.storage_manager": ["get_fns", "llm_load", "llm_save", "init_storage_manager", "get_storage_manager",
                                           "wait_async_upload_finish"],
        "internlm.utils.timeout": ["llm_timeout"],
        "internlm.utils.megatron_timers": ["megatron_timer"],
        "internlm.model.metrics": ["AccPerplex", "SchedulerMetricHook"],
        "internlm.model.losses": ["FlashGPTLMLoss"],
        "internlm.eval.evaluation": ["evaluate_on_val_dls", "switch_evaluation_mode"],
        # deep module paths that scripts written against the reference import directly
        "internlm.core.context.parallel_context": ["Config"],
        "internlm.core.trainer": ["TrainState", "Trainer"],
        "internlm.data.train_state": ["get_train_state"],
        "internlm.initialize.launch": ["args_sanity_check", "launch_from_torch"],
        "internlm.monitor.monitor": ["monitor_manager"],
        "internlm.solver.optimizer.hybrid_zero_optim": ["HybridZeroOptimizer"],
        "internlm.solver.beta2_scheduler": ["Beta2Scheduler"],
        "internlm.solver.lr_scheduler": ["FineTuneCosineAnnealingWarmupLR"],
        "internlm.data.tokenized.dummy_dataset": ["RandomDataset"],
        "internlm.data.tokenized.packed_dataset": ["PackedDatasetWithCut", "PackedDatasetWithoutCuSeqlen"],
        "internlm.data.tokenized.batch_sampler": ["StaticBatchSampler"],
        "internlm.data.tokenized.collaters": ["packed_collate_fn"],
        "internlm.model.ops.linear": ["ScaleColumnParallelLinear", "RewardModelLinear"],
        "internlm.model.ops.norm": ["RMSNorm"],
        "internlm.model.modules.embedding": ["Embedding1D", "RotaryEmbedding"],
        "internlm.model.modules.mlp": ["FeedForward"],
        "internlm.model.modules.multi_head_attention": ["MHA"],
        "internlm.model.utils": ["gather_forward_split_backward", "split_forward_gather_backward"],
        "internlm.utils.gputest": ["empty_cache_and_diag"],
        "internlm.utils.parallel": ["get_parallel_log_file_name"],
        "internlm.core.scheduler.pipeline_scheduler": ["PipelineScheduler", "InterleavedPipelineScheduler"],
        "internlm.core.gradient_handler": ["PipelineSharedModuleGradientHandler"],
    }
    missing = {}
    for mod, names in surface.items():
        m = importlib.import_module(mod)
        miss = [n for n in names if not hasattr(m, n)]
        if miss:
            missing[mod] = miss
    assert not missing, missing


def test_set_parallel_attr_for_param_groups_tags_user_model():
    from internevo_b200 import ops
    from internevo_b200.core.context import IS_REPLICA_ZERO_PARALLEL, IS_TENSOR_ZERO_PARALLEL
    from internevo_b200.train import set_parallel_attr_for_param_groups

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.norm = ops.RMSNorm(8)
            self.ln = ops.LayerNorm(8)
            self.fc = torch.nn.Linear(8, 8)

    m = Tiny()
    set_parallel_attr_for_param_groups(m)
    assert getattr(m.norm.weight, IS_REPLICA_ZERO_PARALLEL) and getattr(m.ln.bias, IS_REPLICA_ZERO_PARALLEL)
    assert getattr(m.fc.weight, IS_TENSOR_ZERO_PARALLEL) and not hasattr(m.fc.weight, IS_REPLICA_ZERO_PARALLEL)
    assert not hasattr(m.norm.weight, IS_TENSOR_ZERO_PARALLEL)
