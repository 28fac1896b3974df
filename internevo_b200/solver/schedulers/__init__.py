from .beta2_scheduler import Beta2Scheduler
from .lr_scheduler import FineTuneCosineAnnealingWarmupLR

__all__ = ["Beta2Scheduler", "FineTuneCosineAnnealingWarmupLR"]
