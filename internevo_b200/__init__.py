"""internevo_b200: a Blackwell (sm_100a) native hybrid-parallel LLM pre-training engine with InternEvo's capabilities."""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not import torch.distributed machinery eagerly
    if name in ("initialize_trainer", "launch_from_slurm", "launch_from_torch", "get_default_parser",
                "initialize_distributed_env"):
        from internevo_b200 import initialize

        return getattr(initialize, name)
    raise AttributeError(name)
