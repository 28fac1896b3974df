"""Runtime diagnostics (reference ``internlm/utils/gputest.py``): process-group warm-up, slow-rank timer diagnosis,
GEMM/attention micro-benchmark, NVLink all-reduce bench, allocator analysis."""
from __future__ import annotations

import gc
import socket

import torch
import torch.distributed as dist

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import get_current_device
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.megatron_timers import megatron_timer as timer

logger = get_logger(__file__)
GLOBAL_PROCESS_GROUP_MODES = [
    ParallelMode.GLOBAL, ParallelMode.DATA, ParallelMode.TENSOR, ParallelMode.PIPELINE, ParallelMode.ZERO1,
    ParallelMode.WEIGHT, ParallelMode.WEIGHT_DATA, ParallelMode.EXPERT, ParallelMode.EXPERT_DATA, ParallelMode.NETTEST,
]
_nccl_retry_seen = 0


def warmup_process_group():
    """One tiny all-reduce per multi-rank group + barrier so communicators are built before step 0 (ref ``:279-302``)."""
    if not gpc.is_distributed:
        return
    dev = get_current_device()
    buf = torch.ones(64, device=dev)
    for mode in GLOBAL_PROCESS_GROUP_MODES:
        group = gpc.get_group(mode)
        if group is not None and gpc.get_world_size(mode) > 1:
            dist.all_reduce(buf, group=group)
    dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def empty_cache_and_diag(batch_count, interval=50):
    """Every ``interval`` steps: slow-rank diagnosis, kernel micro-benchmark, cache flush and explicit GC."""
    if interval <= 0:
        interval = 50
    if batch_count % int(interval) == 0:
        if batch_count > 0:
            if gpc.is_rank_for_log():
                logger.info("Empty Cache and Diagnosis GPU/NCCL/Timer ...")
            with torch.no_grad():
                timer_diagnosis()
                bench_gpu()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        gc.collect()


def benchmark_forward(test_fn, *inputs, repeats: int = 100, amp: bool = True, amp_dtype=torch.float16, warmup: int = 3,
                      **kwinputs) -> float:
    """Mean seconds per call of ``test_fn(*inputs, **kwinputs)`` (forward only, under autocast when ``amp``): CUDA events
    around ``repeats`` back-to-back launches after ``warmup`` untimed ones; the host clock on CPU.  Same contract as the
    reference helper built on ``torch.utils.benchmark`` (``utils/gputest.py:60-80``) without its per-call synchronisation."""
    import time

    on_gpu = torch.cuda.is_available()
    ctx = torch.autocast(device_type="cuda" if on_gpu else "cpu", dtype=amp_dtype if on_gpu else torch.bfloat16, enabled=amp)
    with torch.no_grad(), ctx:
        for _ in range(warmup):
            test_fn(*inputs, **kwinputs)
        if on_gpu:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(repeats):
                test_fn(*inputs, **kwinputs)
            e.record()
            e.synchronize()
            return s.elapsed_time(e) * 1e-3 / repeats
        t0 = time.perf_counter()
        for _ in range(repeats):
            test_fn(*inputs, **kwinputs)
        return (time.perf_counter() - t0) / repeats


def flops(batch, seqlen, headdim, nheads, time_f):
    return (4 * batch * seqlen**2 * nheads * headdim) / time_f / 1e12


def get_gpu_temperature():
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        return pynvml.nvmlDeviceGetTemperature(handle, pynvml.NVML_TEMPERATURE_GPU)
    except Exception:  # pragma: no cover
        return -1


def get_cpu_temperature():
    try:
        import psutil

        temps = psutil.sensors_temperatures()
        return temps["coretemp"][0].current if "coretemp" in temps else -1
    except Exception:  # pragma: no cover
        return -1


def _gather_scalar(value: float, mode: ParallelMode):
    group = gpc.get_group(mode)
    n = gpc.get_world_size(mode)
    if group is None or n <= 1:
        return [value]
    t = torch.tensor([value], device=get_current_device(), dtype=torch.float32)
    out = [torch.zeros_like(t) for _ in range(n)]
    dist.all_gather(out, t, group=group)
    return [float(x) for x in out]


def timer_diagnosis():
    """Compare this rank's named timers against the DP-group (trimmed) mean and its own history; warn on outliers
    beyond ``data.diag_outlier_ratio`` (reference ``:117-178``)."""
    ratio = gpc.config.data.get("diag_outlier_ratio", 1.1) if gpc.config is not None else 1.1
    timer.store_last_timers()
    for name, t in zip(timer.names, timer.times):
        vals = _gather_scalar(t, ParallelMode.DATA)
        if len(vals) > 4:
            vals_sorted = sorted(vals)[1:-1]
        else:
            vals_sorted = vals
        avg = sum(vals_sorted) / max(1, len(vals_sorted))
        if avg > 0 and t > avg * ratio and t - avg > 1e-3:
            logger.warning(f"rank {gpc.get_global_rank()} ({socket.gethostname()}): timer '{name}' {t:.4f}s is "
                           f"{t / avg:.2f}x the data-parallel mean {avg:.4f}s")
        hist = timer(name).history
        if len(hist) >= 3:
            havg = sum(hist) / len(hist)
            if havg > 0 and t > havg * ratio and t - havg > 1e-3:
                logger.warning(f"rank {gpc.get_global_rank()}: timer '{name}' {t:.4f}s vs own history mean {havg:.4f}s")


def bench_net():
    """All-reduce bus bandwidth on the NETTEST group (8 Mi bf16 elements), compared across groups."""
    group = gpc.get_group(ParallelMode.NETTEST)
    n = gpc.get_world_size(ParallelMode.NETTEST)
    if group is None or n <= 1 or not torch.cuda.is_available():
        return None
    buf = torch.ones(8 * 1024 * 1024, device=get_current_device(), dtype=torch.bfloat16)
    for _ in range(2):
        dist.all_reduce(buf, group=group)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        dist.all_reduce(buf, group=group)
    e.record()
    e.synchronize()
    ms = s.elapsed_time(e) / 5
    busbw = buf.numel() * 2 * 2 * (n - 1) / n / (ms * 1e-3) / 1e9
    if gpc.is_rank_for_log():
        logger.info(f"nettest all-reduce: {ms:.3f} ms, bus bandwidth {busbw:.1f} GB/s over {n} ranks")
    return busbw


def bench_gpu(use_flash_attn=True):
    """Micro-benchmark of our own GEMM kernel; warn if this GPU is slower than the global mean (ref ``:227-276``)."""
    if not torch.cuda.is_available():
        return None
    from internevo_b200 import ops

    a = torch.randn(4096, 4096, device=get_current_device(), dtype=torch.bfloat16)
    b = torch.randn(4096, 4096, device=get_current_device(), dtype=torch.bfloat16)
    for _ in range(2):
        ops.matmul(a, b)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        ops.matmul(a, b)
    e.record()
    e.synchronize()
    tfl = 2 * 4096**3 / (s.elapsed_time(e) / 5 * 1e-3) / 1e12
    vals = _gather_scalar(tfl, ParallelMode.GLOBAL)
    mean = sum(vals) / len(vals)
    ratio = gpc.config.data.get("diag_outlier_ratio", 1.1) if gpc.config is not None else 1.1
    if tfl * ratio < mean:
        logger.warning(f"rank {gpc.get_global_rank()} ({socket.gethostname()}) GEMM bench {tfl:.0f} TFLOP/s is below "
                       f"the global mean {mean:.0f}; gpu temp {get_gpu_temperature()}C")
    return tfl


def cuda_memory_analyze(step=0, print_mm_suage=False):
    """Allocator summary + warning when the caching allocator had to retry (reference ``:305-346``)."""
    global _nccl_retry_seen
    if not torch.cuda.is_available():
        return
    g = 1024**3
    stats = torch.cuda.memory_stats()
    retries = stats.get("num_alloc_retries", 0)
    if retries > _nccl_retry_seen:
        _nccl_retry_seen = retries
        logger.warning(f"step {step}: cuda allocator retried {retries} times; memory is nearly exhausted or fragmented")
    if print_mm_suage and gpc.is_rank_for_log():
        logger.info(
            f"step {step}: allocated {torch.cuda.memory_allocated() / g:.2f} GB (max "
            f"{torch.cuda.max_memory_allocated() / g:.2f}), reserved {torch.cuda.memory_reserved() / g:.2f} GB (max "
            f"{torch.cuda.max_memory_reserved() / g:.2f}), frag "
            f"{(torch.cuda.memory_reserved() - torch.cuda.memory_allocated()) / max(1, torch.cuda.memory_reserved()):.2%}"
        )

