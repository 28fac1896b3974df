"""Attribute-style nested config loaded from a python file (same user surface as reference
``internlm/core/context/parallel_context.py:39-127``): every public, non-module, non-class global of the file becomes a
key; nested dicts become nested ``Config`` objects."""
from __future__ import annotations

import importlib.util
import inspect
import os
import sys
from contextlib import contextmanager
from typing import Any, Union


class Config(dict):
    """``dict`` whose items are also attributes; nested dicts are converted recursively."""

    def __init__(self, config: dict = None):
        super().__init__()
        if config is not None:
            for k, v in config.items():
                self._add_item(k, v)

    def __missing__(self, key):
        raise KeyError(key)

    def __getattr__(self, key):
        try:
            return super().__getitem__(key)
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        super().__setitem__(key, value)

    def __setitem__(self, key, value):
        super().__setitem__(key, Config(value) if isinstance(value, dict) and not isinstance(value, Config) else value)

    def _add_item(self, key, value):
        self[key] = value

    def update(self, other=None, **kw):
        for k, v in dict(other or {}, **kw).items():
            self._add_item(k, v)
        return self

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, Config) else v) for k, v in self.items()}

    @staticmethod
    def from_file(filename: str) -> "Config":
        filename = os.path.abspath(filename)
        if not os.path.exists(filename):
            raise FileNotFoundError(f"config file {filename} not found")
        assert filename.endswith(".py"), "only .py config files are supported"
        name = "_b200_cfg_" + os.path.splitext(os.path.basename(filename))[0]
        # the directory that contains ``configs/`` must be importable for ``from configs._base_... import *``
        roots = [os.path.dirname(filename), os.path.dirname(os.path.dirname(filename)), os.getcwd()]
        added = [r for r in roots if r not in sys.path]
        sys.path[:0] = added
        try:
            spec = importlib.util.spec_from_file_location(name, filename)
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
        finally:
            for r in added:
                sys.path.remove(r)
        cfg = Config()
        for k, v in module.__dict__.items():
            if k.startswith("__") or inspect.ismodule(v) or inspect.isclass(v) or inspect.isfunction(v):
                continue
            cfg._add_item(k, v)
        return cfg


@contextmanager
def read_base():
    """No-op context used by configs for ``with read_base(): from configs._base_.x import *`` (reference
    ``internlm/utils/utils.py:5-18``)."""
    yield


def load_config(config: Union[str, dict, Config, Any]) -> Config:
    if isinstance(config, Config):
        return config
    if isinstance(config, dict):
        return Config(config)
    if isinstance(config, (str, os.PathLike)):
        return Config.from_file(str(config))
    raise TypeError(f"config must be a path, dict or Config, got {type(config)}")
