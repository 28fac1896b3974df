from .config import Config, load_config, read_base
from .parallel_context import (
    IS_REPLICA_ZERO_PARALLEL,
    IS_TENSOR_DATA_PARALLEL,
    IS_TENSOR_EXPERT_DATA_PARALLEL,
    IS_TENSOR_ZERO_PARALLEL,
    IS_WEIGHT_ZERO_PARALLEL,
    ParallelContext,
    global_context,
)
from .process_groups import ParallelMode, ParallelSizes, group_rank_lists, layout_for_rank, modes_to_build
from .random import (
    add_seed,
    get_current_mode,
    get_seeds,
    get_states,
    seed,
    set_mode,
    set_seed_states,
    sync_states,
)

__all__ = [
    "Config", "load_config", "read_base", "ParallelContext", "global_context", "ParallelMode", "ParallelSizes",
    "group_rank_lists", "layout_for_rank", "modes_to_build", "IS_REPLICA_ZERO_PARALLEL", "IS_TENSOR_DATA_PARALLEL",
    "IS_TENSOR_EXPERT_DATA_PARALLEL", "IS_TENSOR_ZERO_PARALLEL", "IS_WEIGHT_ZERO_PARALLEL", "add_seed",
    "get_current_mode", "get_seeds", "get_states", "seed", "set_mode", "set_seed_states", "sync_states",
]
