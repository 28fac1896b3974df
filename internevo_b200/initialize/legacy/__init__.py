"""Pre-``load_ckpt_info`` checkpoint config keys (reference ``internlm/initialize/legacy``)."""
from .launch import auto_resume_sanity_check, ckpt_info_sanity_check

__all__ = ["auto_resume_sanity_check", "ckpt_info_sanity_check"]
