from .base_optimizer import BaseOptimizer
from .fsdp_optimizer import FSDPadaptOptimizer
from .hybrid_zero_optim import HybridZeroOptimizer
from .utils import DynamicGradScaler

__all__ = ["BaseOptimizer", "FSDPadaptOptimizer", "HybridZeroOptimizer", "DynamicGradScaler"]
