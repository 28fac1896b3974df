"""Weight init helpers (reference ``internlm/initialize/initialize_tensor.py:9-63``)."""
import math

from torch import Tensor, nn


def scaled_init_method_normal(sigma: float = 1.0, num_layers: int = 1):
    """N(0, sigma / sqrt(2 * num_layers))."""
    std = sigma / math.sqrt(2.0 * num_layers)

    def init_(tensor: Tensor):
        return nn.init.normal_(tensor, mean=0.0, std=std)

    return init_


def normal_(mean: float = 0.0, std: float = 1.0):
    def initializer(tensor: Tensor):
        return nn.init.normal_(tensor, mean, std)

    return initializer


def scaled_init_method_uniform(sigma: float = 1.0, num_layers: int = 1):
    """Uniform(-a, a) parameterised like the reference (a = sqrt(3 * std), std = sigma / sqrt(2 * num_layers))."""
    std = sigma / math.sqrt(2.0 * num_layers)
    a = math.sqrt(3.0 * std)

    def init_(tensor: Tensor):
        return nn.init.uniform_(tensor, -a, a)

    return init_


def uniform_(mean: float = 0.0, std: float = 1.0):
    a = math.sqrt(3.0 * std)

    def initializer(tensor: Tensor):
        return nn.init.uniform_(tensor, mean - a, mean + a)

    return initializer
