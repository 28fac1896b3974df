"""Alert / heartbeat transport (reference ``internlm/monitor/alert.py``): Feishu-compatible webhook POST and a generic
"light monitor" heartbeat endpoint; both are best-effort and never raise into the training loop."""
from __future__ import annotations

import json
import math
import os
import time
from typing import Dict

import requests

from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)
LIGHT_MONITOR_ADDRESS = None


def initialize_light_monitor(monitor_address: str = None):
    global LIGHT_MONITOR_ADDRESS
    LIGHT_MONITOR_ADDRESS = monitor_address


def send_heartbeat(msg_type: str, msg: Dict):
    """POST ``{"type", "time", "job", "rank", "msg"}`` to the light-monitor address, if configured."""
    if not LIGHT_MONITOR_ADDRESS:
        return
    def nan2none(v):
        if isinstance(v, float) and (math.isnan(v) or math.isinf(v)):
            return None
        return v

    payload = {"type": msg_type, "time": time.time(), "job_id": os.environ.get("JOB_NAME", "none"),
               "cluster": os.environ.get("CLUSTER_NAME", "none"), "rank": os.environ.get("RANK", "0"),
               "msg": {k: nan2none(v) for k, v in msg.items() if not isinstance(v, dict)}}
    try:
        requests.post(LIGHT_MONITOR_ADDRESS, data=json.dumps(payload), headers={"Content-Type": "application/json"},
                      timeout=5)
    except Exception as e:  # pragma: no cover
        logger.warning(f"heartbeat failed: {e}")


def send_feishu_msg_with_webhook(webhook: str, title: str, message: str):
    """Feishu/Lark bot 'post' message (reference ``alert.py:90-136``); any webhook accepting that JSON works."""
    headers = {"Content-Type": "application/json;charset=utf-8"}
    msg_body = {"timestamp": int(time.time()), "msg_type": "post",
                "content": {"post": {"zh_cn": {"title": title, "content": [[{"tag": "text", "text": message}]]}}}}
    try:
        res = requests.post(webhook, data=json.dumps(msg_body), headers=headers, timeout=30)
        res = res.json()
        logger.info(f"alert sent, response: {res}")
    except Exception as err:  # pragma: no cover
        logger.error(f"alert send failed: {err}")
