from .fsdp_optimizer import FSDPadaptOptimizer
from .hybrid_zero_optim import HybridZeroOptimizer
from .utils import DynamicGradScaler

__all__ = ["FSDPadaptOptimizer", "HybridZeroOptimizer", "DynamicGradScaler"]
