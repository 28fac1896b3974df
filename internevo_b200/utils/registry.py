"""Name → callable registry (reference ``internlm/utils/registry.py:5-71``)."""
from __future__ import annotations


class Registry:
    def __init__(self, name: str):
        self._name = name
        self._registry = {}

    @property
    def name(self):
        return self._name

    def register_module(self, module_name: str, func=None):
        """Usable as ``@REG.register_module("x")`` or ``REG.register_module("x", fn)``."""
        assert module_name not in self._registry, f"{module_name} already registered in {self._name}"

        def deco(f):
            self._registry[module_name] = f
            return f

        if func is not None:
            return deco(func)
        return deco

    def get_module(self, module_name: str):
        if module_name not in self._registry:
            raise NameError(f"{module_name} not found in registry {self._name}; known: {sorted(self._registry)}")
        return self._registry[module_name]

    def has(self, module_name: str) -> bool:
        return module_name in self._registry

    def keys(self):
        return list(self._registry)


MODEL_INITIALIZER = Registry("model_initializer")
MOE_INITIALIZER = Registry("moe_initializer")
