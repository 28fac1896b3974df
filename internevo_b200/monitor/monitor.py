"""Training monitor (reference ``internlm/monitor/monitor.py``): hang detection thread, loss-spike detection,
exception / SIGTERM alerts, de-duplicated through an flock'd alert file."""
from __future__ import annotations

import fcntl
import os
import signal
import socket
import time
from contextlib import contextmanager
from threading import Thread

from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import SingletonMeta
from internevo_b200.utils.logger import get_logger

from .alert import send_feishu_msg_with_webhook
from .utils import get_job_key, set_env_var

logger = get_logger(__file__)


def send_alert_message(address: str = None, title: str = None, message: str = None):
    """Alert via webhook if an address is configured; always logged on the log rank."""
    if address is not None and gpc.is_rank_for_log():
        send_feishu_msg_with_webhook(webhook=address, title=title if title else get_job_key(), message=message)
    elif gpc.is_rank_for_log() and message:
        logger.warning(f"[alert] {title or get_job_key()}: {message}")


def _env_number(key: str):
    """Numeric value another part of the process published through the environment, or ``None``."""
    raw = os.getenv(key)
    if raw is None:
        return None
    try:
        return float(raw)
    except ValueError:
        return None


class MonitorTracker(Thread):
    """Watchdog thread.  The training loop publishes a heartbeat (``LAST_ACTIVE_TIMESTAMP``) and, on the log rank, the
    latest ``LOSS`` / ``STEP_ID`` through the process environment (``train/pipeline.py``); every ``check_interval`` seconds
    this thread compares them with what it saw last time: a heartbeat that did not advance means a hang, a loss that grew by
    more than ``loss_spike_limit`` x means a spike (reference behaviour: ``internlm/monitor/monitor.py:35-118``)."""

    daemon = True

    def __init__(self, alert_address: str, check_interval: float = 300, loss_spike_limit: float = 1.5):
        super().__init__()
        self.alert_address = alert_address
        self.check_interval = check_interval
        self.loss_spike_limit = loss_spike_limit
        self._seen_heartbeat = None
        self._seen_loss = None
        self.stopped = False
        self.start()

    def run(self):
        while not self.stopped:
            for probe in (self._check_stuck, self._check_loss_spike):
                try:
                    probe()
                except Exception as e:  # pragma: no cover - a monitoring failure must never take the job down
                    logger.debug(f"monitor probe failed: {e}")
            deadline = time.monotonic() + self.check_interval
            while not self.stopped and time.monotonic() < deadline:
                time.sleep(min(1.0, max(0.0, deadline - time.monotonic())))

    def _check_stuck(self):
        beat = _env_number("LAST_ACTIVE_TIMESTAMP")
        if beat is not None and self._seen_heartbeat is not None and beat <= self._seen_heartbeat:
            self._send_alert(f"No training step finished during the last {self.check_interval:.0f} s: the job may be stuck.")
        if beat is not None:
            self._seen_heartbeat = beat

    def _check_loss_spike(self):
        if not gpc.is_rank_for_log():
            return
        loss, step = _env_number("LOSS"), _env_number("STEP_ID")
        if loss is None:
            return
        prev, self._seen_loss = self._seen_loss, loss
        if prev is not None and prev > 0 and loss / prev > self.loss_spike_limit:
            self._send_alert(f"Periodic check: possible loss spike around step {int(step) if step is not None else '?'} "
                             f"({prev:.4f} -> {loss:.4f}).")

    def _send_alert(self, message):
        send_alert_message(address=self.alert_address, message=message)

    def stop(self):
        self.stopped = True


class MonitorManager(metaclass=SingletonMeta):
    """Process-wide switchboard: owns the watchdog thread, per-step spike check, exception and SIGTERM alerts."""

    def __init__(self, loss_spike_limit: float = 1.5) -> None:
        self.monitor_thread = None
        self.loss_spike_limit = loss_spike_limit
        self.last_step_loss = -1
        self.alert_file_path = None
        self.enable_alert = False
        self.light_monitor_address = None

    def monitor_loss_spike(self, alert_address: str = None, step_count: int = 0, cur_step_loss: float = 0.0):
        """Per-step check (the watchdog repeats it at its own period); also publishes loss / step for the watchdog."""
        if not self.enable_alert:
            return
        set_env_var(key="LOSS", value=cur_step_loss)
        set_env_var(key="STEP_ID", value=step_count)
        prev, self.last_step_loss = self.last_step_loss, cur_step_loss
        if prev != -1 and cur_step_loss > self.loss_spike_limit * prev:
            send_alert_message(address=alert_address,
                               message=f"Step {step_count}: loss went from {prev} to {cur_step_loss} "
                                       f"(> {self.loss_spike_limit}x), possible loss spike.")

    def exception_should_be_alert(self, msg: str, alert_address: str = None):
        """De-duplication across ranks: the alert file is an flock'd set of already reported messages; only the rank that
        adds a message reports it (reference ``:158-176``)."""
        if not self.enable_alert:
            return False
        if self.alert_file_path is None:
            return True
        try:
            os.makedirs(os.path.dirname(self.alert_file_path) or ".", exist_ok=True)
            with open(self.alert_file_path, "a+") as f:
                fcntl.flock(f, fcntl.LOCK_EX)
                try:
                    f.seek(0)
                    known = set(f.read().splitlines())
                    if msg in known:
                        return False
                    f.write(msg + "\n")
                    return True
                finally:
                    fcntl.flock(f, fcntl.LOCK_UN)
        except OSError:  # pragma: no cover - an unwritable alert file must not hide the exception
            return True

    def monitor_exception(self, alert_address: str = None, excp_info: str = None):
        if not self.enable_alert:
            return
        tail = (excp_info or "").strip().split("\n")[-10:]
        if not self.exception_should_be_alert(tail[-1] if tail else "", alert_address):
            return
        message = (f"Exception on {socket.gethostname()} (global rank {gpc.get_global_rank()}):\n" + "\n".join(tail))
        # exceptions are reported by whichever rank hits them, not only the log rank
        if alert_address:
            send_feishu_msg_with_webhook(alert_address, get_job_key(), message)
        else:
            logger.error(message)

    def handle_sigterm(self, alert_address: str = None):
        def on_sigterm(signum, frame):
            send_alert_message(address=alert_address,
                               message=f"Process on {socket.gethostname()} received signal {signum} and is exiting.")
            raise SystemExit(128 + signum)

        signal.signal(signal.SIGTERM, on_sigterm)

    def start_monitor(self, job_name: str, alert_address: str, monitor_interval_seconds: int = 300,
                      loss_spike_limit: float = 1.5):
        set_env_var(key="JOB_NAME", value=job_name)
        self.enable_alert = True
        self.loss_spike_limit = loss_spike_limit
        self.monitor_thread = MonitorTracker(alert_address=alert_address, check_interval=monitor_interval_seconds,
                                             loss_spike_limit=loss_spike_limit)

    def stop_monitor(self):
        if self.monitor_thread is not None:
            self.monitor_thread.stop()
            self.monitor_thread = None
        self.enable_alert = False


monitor_manager = MonitorManager()


@contextmanager
def initialize_monitor_manager(job_name: str = None, alert_address: str = None):
    """``with initialize_monitor_manager(job_name, address): main()`` — starts the tracker when an alert address is set,
    reports exceptions, always stops the thread (reference ``:265-300``)."""
    if alert_address is not None:
        try:
            monitor_manager.start_monitor(job_name=job_name, alert_address=alert_address)
            monitor_manager.handle_sigterm(alert_address=alert_address)
            send_alert_message(address=alert_address, message=f"Training in {socket.gethostname()} is starting.")
            yield
        finally:
            send_alert_message(address=alert_address, message=f"Training in {socket.gethostname()} completed.")
            monitor_manager.stop_monitor()
    else:
        yield
