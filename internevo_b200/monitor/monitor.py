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


class MonitorTracker(Thread):
    """Polls ``LAST_ACTIVE_TIMESTAMP`` (refreshed every step) and alerts when training looks stuck."""

    def __init__(self, alert_address: str, check_interval: float = 300, loss_spike_limit: float = 1.5):
        super().__init__()
        self.alert_address = alert_address
        self.check_interval = check_interval
        self.loss_spike_limit = loss_spike_limit
        self.last_active_time = -1
        self.last_loss_value = -1
        self.stopped = False
        self.start()

    def run(self):
        while not self.stopped:
            try:
                self._check_stuck()
                self._check_loss_spike()
            except Exception:  # pragma: no cover
                continue
            slept = 0.0
            while slept < self.check_interval and not self.stopped:
                time.sleep(min(1.0, self.check_interval - slept))
                slept += 1.0

    def _check_stuck(self):
        new_active_time = -1
        if os.getenv("LAST_ACTIVE_TIMESTAMP") is not None:
            new_active_time = os.getenv("LAST_ACTIVE_TIMESTAMP")
        if int(new_active_time) <= int(self.last_active_time) and new_active_time != -1:
            self._send_alert("Training may be in stuck status, please check it.")
        self.last_active_time = new_active_time

    def _check_loss_spike(self):
        if gpc.is_rank_for_log():
            new_loss_value = -1
            new_step_id = -1
            if os.getenv("LOSS") is not None:
                new_loss_value = os.getenv("LOSS")
            if os.getenv("STEP_ID") is not None:
                new_step_id = os.getenv("STEP_ID")
            if (float(new_loss_value) / float(self.last_loss_value)) > self.loss_spike_limit and new_loss_value != -1:
                assert int(new_step_id) >= 0
                self._send_alert(f"Checking periodically: Loss spike may be happened in step {new_step_id}, "
                                 f"loss value from {self.last_loss_value} to {new_loss_value}, please check it.")
            self.last_loss_value = new_loss_value

    def _send_alert(self, message):
        send_alert_message(address=self.alert_address, message=message)

    def stop(self):
        self.stopped = True


class MonitorManager(metaclass=SingletonMeta):
    def __init__(self, loss_spike_limit: float = 1.5) -> None:
        self.monitor_thread = None
        self.loss_spike_limit = loss_spike_limit
        self.last_step_loss = -1
        self.alert_file_path = None
        self.enable_alert = False
        self.light_monitor_address = None

    def monitor_loss_spike(self, alert_address: str = None, step_count: int = 0, cur_step_loss: float = 0.0):
        """Alert when the loss jumps by more than ``loss_spike_limit`` between consecutive steps."""
        if self.enable_alert:
            set_env_var(key="LOSS", value=cur_step_loss)
            set_env_var(key="STEP_ID", value=step_count)
            if self.last_step_loss != -1 and cur_step_loss > self.loss_spike_limit * self.last_step_loss:
                send_alert_message(address=alert_address, message=(
                    f"Checking step by step: Loss spike may be happened in step {step_count}, "
                    f"loss value from {self.last_step_loss} to {cur_step_loss}, please check it."))
            self.last_step_loss = cur_step_loss

    def exception_should_be_alert(self, msg: str, alert_address: str = None):
        """Only the first rank to write the (flock'd) alert file sends; the rest stay quiet (reference ``:158-176``)."""
        if self.enable_alert is False:
            return False
        if self.alert_file_path is None:
            return True
        try:
            os.makedirs(os.path.dirname(self.alert_file_path) or ".", exist_ok=True)
            with open(self.alert_file_path, "a+") as f:
                fcntl.flock(f, fcntl.LOCK_EX)
                f.seek(0)
                if msg in f.read():
                    fcntl.flock(f, fcntl.LOCK_UN)
                    return False
                f.write(msg + "\n")
                fcntl.flock(f, fcntl.LOCK_UN)
            return True
        except Exception:  # pragma: no cover
            return True

    def monitor_exception(self, alert_address: str = None, excp_info: str = None):
        if self.enable_alert:
            filtered = excp_info.split("\n")[-10:]
            msg = "\n".join(filtered)
            if self.exception_should_be_alert(filtered[-1] if filtered else "", alert_address):
                message = f"Catch Exception from {socket.gethostname()} with rank id {gpc.get_global_rank()}:{msg}"
                # exceptions are reported by whichever rank hits them, not only the log rank
                if alert_address:
                    send_feishu_msg_with_webhook(alert_address, get_job_key(), message)
                else:
                    logger.error(message)

    def handle_sigterm(self, alert_address: str = None):
        def sigterm_handler(sys_signal, frame):
            message = f"Process received signal {signal} and exited."
            send_alert_message(address=alert_address, message=message)
            raise SystemExit(128 + sys_signal)

        signal.signal(signal.SIGTERM, sigterm_handler)

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
