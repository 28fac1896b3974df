"""Job identity for alerts and heartbeats (counterpart of the reference's ``internlm/monitor/utils.py``).

A job is named ``<scheduler id>_<JOB_NAME>``.  The scheduler id is looked up in a table of (environment variable, extractor)
pairs - Slurm, Kubernetes pods (the pod-name prefix), the Volc MLP platform - so another scheduler is one more table row.
"""
import importlib
import os
import time
from typing import Callable, Optional, Tuple

# (probe variable, how to turn the environment into an id); first hit wins
_JOB_ID_SOURCES: Tuple[Tuple[str, Callable[[], Optional[str]]], ...] = (
    ("SLURM_JOB_ID", lambda: os.environ["SLURM_JOB_ID"]),
    ("K8S_WORKSPACE_ID", lambda: os.environ["K8S_WORKSPACE_ID"]),
    ("KUBERNETES_POD_NAME", lambda: os.environ["KUBERNETES_POD_NAME"].split("-", 1)[0]),
    ("MLP_TASK_INSTANCE_ID", lambda: os.environ.get("MLP_TASK_ID")),
)


def now_time() -> str:
    """Wall-clock stamp used in alert texts."""
    return time.strftime("%Y-%m-%d %H:%M:%S", time.localtime())


def set_env_var(key, value) -> None:
    os.environ[str(key)] = str(value)


def get_job_id() -> str:
    for probe, extract in _JOB_ID_SOURCES:
        if probe in os.environ:
            found = extract()
            if found:
                return found
    return "none"


def get_job_name() -> str:
    return os.environ.get("JOB_NAME") or "unknown"


def get_job_key() -> str:
    return "_".join((get_job_id(), get_job_name()))


def try_import_send_exception():
    """The optional site-specific exception reporter: ``uniscale_monitoring.send_exception_msg`` when that (proprietary)
    package is importable, else ``None`` and the caller falls back to the webhook alert (reference ``monitor/utils.py:37-47``)."""
    try:
        return getattr(importlib.import_module("uniscale_monitoring"), "send_exception_msg", None)
    except ImportError:
        return None
