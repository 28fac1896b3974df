"""Initialiser factories used by the model builders: each returns a callable ``tensor -> tensor`` (in place).

Names and numerics follow the reference (``internlm/initialize/initialize_tensor.py:9-63``) because configs and the
depth-scaled init of the output projections refer to them; all four are views of one parameterised factory.
"""
import math
from typing import Callable

from torch import Tensor, nn

Init = Callable[[Tensor], Tensor]


def _make(kind: str, mean: float, std: float) -> Init:
    if kind == "normal":
        return lambda t: nn.init.normal_(t, mean=mean, std=std)
    half_width = math.sqrt(3.0 * std)      # the reference's (unusual) width: sqrt(3 * std), kept for init parity
    return lambda t: nn.init.uniform_(t, mean - half_width, mean + half_width)


def _depth_scaled(sigma: float, num_layers: int) -> float:
    """Std of residual-branch output projections: ``sigma / sqrt(2 * depth)``."""
    return sigma / math.sqrt(2.0 * num_layers)


def normal_(mean: float = 0.0, std: float = 1.0) -> Init:
    return _make("normal", mean, std)


def uniform_(mean: float = 0.0, std: float = 1.0) -> Init:
    return _make("uniform", mean, std)


def scaled_init_method_normal(sigma: float = 1.0, num_layers: int = 1) -> Init:
    return _make("normal", 0.0, _depth_scaled(sigma, num_layers))


def scaled_init_method_uniform(sigma: float = 1.0, num_layers: int = 1) -> Init:
    return _make("uniform", 0.0, _depth_scaled(sigma, num_layers))
