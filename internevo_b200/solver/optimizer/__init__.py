from .hybrid_zero_optim import HybridZeroOptimizer
from .utils import DynamicGradScaler

__all__ = ["HybridZeroOptimizer", "DynamicGradScaler"]
