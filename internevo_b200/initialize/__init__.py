from .initialize_trainer import initialize_trainer
from .launch import (
    args_sanity_check,
    get_default_parser,
    initialize_distributed_env,
    launch_from_slurm,
    launch_from_torch,
    try_bind_numa,
)

__all__ = ["get_default_parser", "initialize_distributed_env", "launch_from_slurm", "launch_from_torch",
           "initialize_trainer", "try_bind_numa", "args_sanity_check"]
