"""Process-wide logger (reference ``internlm/utils/logger.py:18-98``): one named logger with a stream handler and an
optional per-rank file handler; the proprietary ``uniscale`` sink is replaced by a pluggable handler list."""
from __future__ import annotations

import logging
import os

LOGGER_NAME = "internevo_b200"
LOGGER_FORMAT = "%(asctime)s\t%(levelname)s %(filename)s:%(lineno)s in %(funcName)s -- %(message)s"
LOGGER_LEVEL = os.environ.get("INTERNEVO_LOG_LEVEL", "info")
_LEVELS = {"info": logging.INFO, "debug": logging.DEBUG, "warning": logging.WARNING, "error": logging.ERROR}
_extra_handlers = []


def get_logger(logger_name: str = LOGGER_NAME, logging_level: str = LOGGER_LEVEL) -> logging.Logger:
    logger = logging.getLogger(LOGGER_NAME)
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(LOGGER_FORMAT))
        logger.addHandler(handler)
        logger.propagate = False
    logger.setLevel(_LEVELS.get(str(logging_level).lower(), logging.INFO))
    return logger


def add_handler(handler: logging.Handler) -> None:
    """Attach an additional sink (e.g. a monitoring platform)."""
    _extra_handlers.append(handler)
    get_logger().addHandler(handler)


def initialize_uniscale_logger(job_name=None, launch_time=None, file_name=None, name=LOGGER_NAME, level=LOGGER_LEVEL,
                               file_path=None, is_std=True):
    """Per-rank file logging under ``RUN/{job}/{time}/logs`` (same directory convention as the reference)."""
    logger = get_logger(name, level)
    if job_name and launch_time and file_name:
        log_dir = file_path or os.path.join("RUN", job_name, launch_time, "logs")
        os.makedirs(log_dir, exist_ok=True)
        fh = logging.FileHandler(os.path.join(log_dir, f"{file_name}.log"))
        fh.setFormatter(logging.Formatter(LOGGER_FORMAT))
        logger.addHandler(fh)
    return logger
