"""Host-side utilities: storage manager (local backend, async uploads + completion marker, custom backends), timeout
decorator, timers, grad scaler, LR / beta2 schedules, registry, monitor (loss-spike / exception filter), config loader
(reference: tests/test_utils/test_storage_manager.py, test_timeout.py and the solver tests)."""
import math
import os
import time

import pytest
import torch


def test_storage_local_and_async(tmp_path):
    from internevo_b200.utils import storage_manager as sm

    assert sm.try_get_storage_backend("boto3:s3://bucket.ep/x/y") == ("boto3", "s3://bucket.ep/x/y")
    assert sm.try_get_storage_backend("local:/a/b") == ("local", "/a/b")
    mgr = sm.init_storage_manager(True, None, False)
    p = f"local:{tmp_path}/a/b/obj.pt"
    sm.llm_save(p, {"x": torch.arange(4)})
    assert torch.equal(sm.llm_load(p)["x"], torch.arange(4))
    assert sm.get_fns(f"local:{tmp_path}/a/b") == ["obj.pt"]
    sm.check_folder(f"local:{tmp_path}/a/b")
    with pytest.raises(AssertionError):
        sm.check_folder(f"local:{tmp_path}/missing")
    mgr.delete_obj(p)
    assert sm.get_fns(f"local:{tmp_path}/a/b") == []

    # a custom object-store backend: uploads go through the thread pool, the marker appears only after wait()
    store = {}

    class Mem(sm.StorageClient):
        def upload_file(self, local_path, remote):
            time.sleep(0.05)
            store[remote] = open(local_path, "rb").read()

        def upload_bytes(self, data, remote):
            store[remote] = data

        def download_bytes(self, remote):
            return store[remote]

        def list(self, remote):
            return sorted({k[len(remote):].lstrip("/").split("/")[0] for k in store if k.startswith(remote)})

        def exists(self, remote):
            return any(k.startswith(remote) for k in store)

        def delete(self, remote):
            store.pop(remote, None)

    sm.register_backend("volc", lambda path: Mem())
    mgr = sm.init_storage_manager(True, str(tmp_path / "staging"), True)
    for i in range(3):
        sm.llm_save(f"volc:vc://bkt/ckpt/10/part{i}.pt", {"i": i})
    mgr.set_pending_marker("volc:vc://bkt/ckpt/10/10.step")
    assert "vc://bkt/ckpt/10/10.step" not in store
    assert mgr.wait()
    assert "vc://bkt/ckpt/10/10.step" in store and "vc://bkt/ckpt/10/part2.pt.md5" in store
    assert sm.llm_load("volc:vc://bkt/ckpt/10/part1.pt")["i"] == 1
    assert os.listdir(tmp_path / "staging") == []   # staging files are removed after upload
    sm.init_storage_manager(False, None, False)


def test_llm_timeout_raises():
    from internevo_b200.utils.timeout import Timeout, llm_timeout

    @llm_timeout(seconds=1, func_name="sleepy")
    def sleepy(t):
        time.sleep(t)
        return "done"

    assert sleepy(0.01) == "done"
    with pytest.raises(TimeoutError):
        sleepy(3)
    with Timeout(2, "inner"):
        time.sleep(0.01)


def test_timers_accumulate_and_history():
    from internevo_b200.utils.megatron_timers import Timers

    timers = Timers()
    for _ in range(3):
        timers("fwd").start()
        time.sleep(0.01)
        timers("fwd").stop()
    e = timers("fwd").elapsed(reset=True)
    assert 0.02 < e < 0.5
    assert timers("fwd").elapsed(reset=False) == 0


def test_dynamic_grad_scaler():
    from internevo_b200.solver.optimizer.utils import DynamicGradScaler

    s = DynamicGradScaler(initial_scale=2 ** 10, growth_factor=2, backoff_factor=0.5, growth_interval=3, min_scale=1,
                          max_scale=2 ** 12, hysteresis=2)
    s.update(True)
    assert s.scale == 2 ** 10          # hysteresis: first overflow is tolerated
    s.update(True)
    assert s.scale == 2 ** 9
    for _ in range(3):
        s.update(False)
    assert s.scale == 2 ** 10
    for _ in range(30):
        s.update(False)
    assert s.scale == 2 ** 12          # clamped at max_scale
    t = DynamicGradScaler()
    t.load_state_dict(s.state_dict())
    assert t.scale == s.scale


def test_lr_and_beta2_schedules():
    from internevo_b200.solver.schedulers import Beta2Scheduler, FineTuneCosineAnnealingWarmupLR

    class Opt:
        param_groups = [{"lr": 1.0, "betas": (0.9, 0.95)}]

    opt = Opt()
    sch = FineTuneCosineAnnealingWarmupLR(opt, total_steps=100, init_steps=2, warmup_ratio=0.1, eta_min=0.1)
    lrs = [opt.param_groups[0]["lr"]]
    for _ in range(99):
        sch.step()
        lrs.append(opt.param_groups[0]["lr"])
    assert lrs[0] == 0 and lrs[1] == 0                      # init_steps
    assert lrs[2] == pytest.approx(0.1) and lrs[11] == pytest.approx(1.0)   # linear warm-up over 10 steps
    assert lrs[12] == pytest.approx(1.0) and all(a >= b - 1e-12 for a, b in zip(lrs[12:], lrs[13:]))
    mid = 12 + (100 - 12) // 2
    assert lrs[mid] == pytest.approx(0.1 + 0.9 * (1 + math.cos(math.pi * (mid - 12) / 88)) / 2)
    b2 = Beta2Scheduler(opt, init_beta2=0.95, c=0.8, cur_iter=-1)
    for _ in range(50):
        b2.step()
    assert opt.param_groups[0]["betas"][1] == pytest.approx(max(0.95, 1 - 1 / 50 ** 0.8))


def test_registry_and_config(tmp_path):
    from internevo_b200.core.context import Config
    from internevo_b200.utils.registry import Registry

    reg = Registry("things")

    @reg.register_module("a")
    def build_a():
        return "A"

    assert reg.get_module("a")() == "A" and reg.has("a")
    with pytest.raises(AssertionError):
        reg.register_module("a")(build_a)
    with pytest.raises(Exception):
        reg.get_module("missing")
    (tmp_path / "base.py").write_text("HIDDEN = 64\nmodel = dict(hidden_size=HIDDEN, num_layers=2)\n")
    (tmp_path / "cfg.py").write_text(
        "from internevo_b200.utils.common import read_base\nwith read_base():\n    from base import *  # noqa\n"
        "model['num_layers'] = 4\ndata = dict(seq_len=HIDDEN * 2)\nimport os\n")
    import sys

    sys.path.insert(0, str(tmp_path))
    try:
        cfg = Config.from_file(str(tmp_path / "cfg.py"))
    finally:
        sys.path.remove(str(tmp_path))
    assert cfg.model.hidden_size == 64 and cfg.model.num_layers == 4 and cfg.data.seq_len == 128
    assert "os" not in cfg      # modules are not copied into the config
    cfg.model._add_item("extra", 1)
    assert cfg.model.extra == 1


def test_monitor_loss_spike_and_exception_filter(tmp_path):
    from internevo_b200.monitor import monitor as mon

    sent = []
    orig = mon.send_alert_message
    mon.send_alert_message = lambda address=None, title=None, message=None: sent.append(message)
    try:
        mm = mon.MonitorManager(loss_spike_limit=1.5)
        mm.enable_alert = True   # as set by start_monitor()
        mm.cur_step_loss = -1.0
        mm.last_step_loss = -1.0
        mm.monitor_loss_spike(alert_address="x", step_count=1, cur_step_loss=2.0)
        mm.monitor_loss_spike(alert_address="x", step_count=2, cur_step_loss=2.1)
        assert not sent
        mm.monitor_loss_spike(alert_address="x", step_count=3, cur_step_loss=4.0)
        assert sent and "spike" in sent[-1].lower()
        # exceptions: every rank may hit one, only the first writer of the (flock'd) alert file reports it
        hooks = []
        orig_hook = mon.send_feishu_msg_with_webhook
        mon.send_feishu_msg_with_webhook = lambda addr, title, msg: hooks.append(msg)
        mm.alert_file_path = str(tmp_path / "alerts" / "job_alert.log")
        mm.monitor_exception(alert_address="x", excp_info="Traceback ...\nRuntimeError: boom")
        mm.monitor_exception(alert_address="x", excp_info="Traceback ...\nRuntimeError: boom")
        mm.monitor_exception(alert_address="x", excp_info="Traceback ...\nValueError: other")
        mon.send_feishu_msg_with_webhook = orig_hook
        assert len(hooks) == 2 and "boom" in hooks[0] and "other" in hooks[1]
    finally:
        mon.send_alert_message = orig


def test_layernorm_module_cpu_matches_torch():
    """ops.LayerNorm (CPU / fp32 path): same numbers as nn.LayerNorm, block-prologue calling convention."""
    import torch

    from internevo_b200 import ops

    torch.manual_seed(0)
    ln, ref = ops.LayerNorm(32, eps=1e-5), torch.nn.LayerNorm(32, eps=1e-5)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_()
        ref.weight.copy_(ln.weight)
        ref.bias.copy_(ln.bias)
    x, r = torch.randn(5, 32, requires_grad=True), torch.randn(5, 32)
    y, nr = ln(x, r)
    assert torch.allclose(y, ref(x + r), atol=1e-5) and torch.allclose(nr, x + r)
    y.sum().backward()
    assert x.grad is not None and ln.weight.grad is not None and ln.bias.grad is not None
    assert torch.allclose(ln(x), ref(x), atol=1e-5)
    assert set(ln.state_dict()) == {"weight", "bias"}


def test_monitor_watchdog_detects_a_stalled_heartbeat_and_periodic_loss_spike(monkeypatch):
    """The tracker thread's two probes, driven by hand: a heartbeat that does not advance between two checks raises the
    "stuck" alert, a loss that grows past the limit between two checks raises the spike alert."""
    from internevo_b200.monitor import monitor as mon

    sent = []
    monkeypatch.setattr(mon, "send_alert_message", lambda address=None, title=None, message=None: sent.append(message))
    monkeypatch.setattr(mon.MonitorTracker, "start", lambda self: None)   # no thread: call the probes directly
    t = mon.MonitorTracker(alert_address="x", check_interval=5, loss_spike_limit=1.5)
    monkeypatch.delenv("LAST_ACTIVE_TIMESTAMP", raising=False)
    t._check_stuck()
    assert not sent                      # no heartbeat published yet: nothing to compare
    monkeypatch.setenv("LAST_ACTIVE_TIMESTAMP", "100")
    t._check_stuck()
    monkeypatch.setenv("LAST_ACTIVE_TIMESTAMP", "130")
    t._check_stuck()
    assert not sent                      # heartbeat advanced
    t._check_stuck()
    assert len(sent) == 1 and "stuck" in sent[0]
    monkeypatch.setenv("LOSS", "2.0")
    monkeypatch.setenv("STEP_ID", "7")
    t._check_loss_spike()
    monkeypatch.setenv("LOSS", "2.5")
    t._check_loss_spike()
    assert len(sent) == 1
    monkeypatch.setenv("LOSS", "5.0")
    monkeypatch.setenv("STEP_ID", "9")
    t._check_loss_spike()
    assert len(sent) == 2 and "step 9" in sent[1]
