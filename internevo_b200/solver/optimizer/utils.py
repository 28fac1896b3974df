"""Loss scaling (reference ``internlm/solver/optimizer/utils.py:381-543``)."""
from __future__ import annotations

from typing import Optional

from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)


class DynamicGradScaler:
    """Dynamic loss scale with growth interval / backoff / hysteresis. State is kept on the host (a python float): the
    optimizer reads the overflow flag once per step together with the grad norm, so no extra device sync is needed."""

    def __init__(self, initial_scale: float = 2**16, growth_factor: float = 2, backoff_factor: float = 0.5,
                 growth_interval: int = 1000, min_scale: Optional[float] = None, max_scale: Optional[float] = None,
                 hysteresis: int = 2):
        assert growth_factor > 1 and 0 < backoff_factor < 1 and hysteresis >= 0
        self._scale = float(initial_scale)
        self._min_scale, self._max_scale = min_scale, max_scale
        self._growth_factor, self._backoff_factor = growth_factor, backoff_factor
        self._growth_interval = growth_interval
        self._growth_step = 0
        self._hysteresis = hysteresis
        self._hysteresis_step = 0

    @property
    def scale(self) -> float:
        return self._scale

    @property
    def inv_scale(self) -> float:
        return 1.0 / self._scale

    def update(self, overflow: bool) -> None:
        if overflow:
            self._hysteresis_step += 1
            self._growth_step = 0
            if self._hysteresis_step >= self._hysteresis:
                self._scale *= self._backoff_factor
                if self._min_scale:
                    self._scale = max(self._scale, self._min_scale)
                if gpc.is_rank_for_log():
                    logger.warning(f"Overflow occurs, the loss scale is adjusted to {self._scale}")
        else:
            self._growth_step += 1
            if self._growth_step == self._growth_interval:
                self._growth_step = 0
                self._hysteresis_step = 0
                self._scale *= self._growth_factor
                if self._max_scale:
                    self._scale = min(self._scale, self._max_scale)
                if gpc.is_rank_for_log():
                    logger.warning(f"No overflow for consecutive {self._growth_interval} steps, "
                                   f"the loss scale is adjusted to {self._scale}")

    def state_dict(self):
        return {"_scale": self._scale, "_growth_step": self._growth_step, "_hysteresis_step": self._hysteresis_step}

    def load_state_dict(self, state_dict):
        self._scale = float(state_dict["_scale"])
        self._growth_step = state_dict["_growth_step"]
        self._hysteresis_step = state_dict["_hysteresis_step"]
