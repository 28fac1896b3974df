"""Environment helpers shared by the monitor (reference ``internlm/monitor/utils.py``)."""
import os
from datetime import datetime


def now_time():
    return datetime.now().strftime("%Y-%m-%d %H:%M:%S")


def set_env_var(key, value):
    os.environ[str(key)] = str(value)


def get_job_id():
    job_id = "none"
    if os.getenv("SLURM_JOB_ID") is not None:
        job_id = os.getenv("SLURM_JOB_ID")
    elif os.getenv("K8S_WORKSPACE_ID") is not None:
        job_id = os.getenv("K8S_WORKSPACE_ID")
    return job_id


def get_job_name():
    return os.getenv("JOB_NAME", "unknown")


def get_job_key():
    return f"{get_job_id()}_{get_job_name()}"


def try_import_send_exception():
    """The optional site-specific exception reporter: ``uniscale_monitoring.send_exception_msg`` when that (proprietary)
    package is importable, else ``None`` and the caller falls back to the webhook alert (reference ``monitor/utils.py:37-47``)."""
    import importlib

    try:
        return getattr(importlib.import_module("uniscale_monitoring"), "send_exception_msg", None)
    except ImportError:
        return None
