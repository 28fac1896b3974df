from .optimizer import HybridZeroOptimizer
from .schedulers import Beta2Scheduler, FineTuneCosineAnnealingWarmupLR

__all__ = ["Beta2Scheduler", "FineTuneCosineAnnealingWarmupLR", "HybridZeroOptimizer"]
