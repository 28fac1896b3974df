from .pipeline import (
    get_scheduler_hooks,
    initialize_isp_communicator,
    initialize_llm_profile,
    initialize_model,
    initialize_optimizer,
    load_new_batch,
    record_current_batch_training_metrics,
    set_fp32_attr_for_model,
    set_parallel_attr_for_param_groups,
    wrap_FSDP_model,
)
from .utils import create_param_groups

__all__ = ["initialize_llm_profile", "initialize_model", "initialize_isp_communicator", "initialize_optimizer",
           "load_new_batch", "record_current_batch_training_metrics", "get_scheduler_hooks", "create_param_groups",
           "set_fp32_attr_for_model", "set_parallel_attr_for_param_groups", "wrap_FSDP_model"]
