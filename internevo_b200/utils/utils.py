"""Small helpers with the reference's import path (``internlm/utils/utils.py``): ``read_base`` for config composition."""
from internevo_b200.core.context.config import read_base  # noqa: F401

__all__ = ["read_base"]
