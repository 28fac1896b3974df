"""Translate the old checkpoint config surface — ``load_ckpt_folder`` / ``load_model_only_folder`` / ``load_optimizer`` /
``load_given_ckpt`` — into the ``load_ckpt_info`` / ``auto_resume`` form the CheckpointManager works with
(reference ``internlm/initialize/legacy/launch.py:10-44``)."""
from __future__ import annotations


def _get(cfg, key, default):
    try:
        v = cfg.get(key, default) if hasattr(cfg, "get") else getattr(cfg, key, default)
    except (KeyError, AttributeError):
        v = default
    return default if v is None else v


def auto_resume_sanity_check(ckpt_config) -> bool:
    """Old configs said ``load_given_ckpt=True`` to mean "use the folder I named, do not look for a newer snapshot"."""
    given = _get(ckpt_config, "load_given_ckpt", None)
    return True if given is None else not given


def ckpt_info_sanity_check(ckpt_config):
    """→ a ``load_ckpt_info`` dict, or None when the old keys name nothing to load."""
    full, model_only = _get(ckpt_config, "load_ckpt_folder", None), _get(ckpt_config, "load_model_only_folder", None)
    if model_only is not None:
        assert full is None, "'load_ckpt_folder' and 'load_model_only_folder' cannot both be set"
        return dict(path=model_only, content=("model",), ckpt_type="internlm")
    if full is None:
        return None
    assert isinstance(full, str), f"unsupported type {type(full)} for ckpt.load_ckpt_folder"
    content = ("model", "sampler", "optimizer") if _get(ckpt_config, "load_optimizer", True) else ("model", "sampler")
    return dict(path=full, content=content, ckpt_type="internlm")
