from .alert import initialize_light_monitor, send_feishu_msg_with_webhook, send_heartbeat
from .monitor import initialize_monitor_manager, monitor_manager, send_alert_message
from .utils import set_env_var

__all__ = ["send_alert_message", "initialize_monitor_manager", "send_feishu_msg_with_webhook", "set_env_var",
           "send_heartbeat", "initialize_light_monitor", "monitor_manager"]
