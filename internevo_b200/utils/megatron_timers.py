"""Named timers. The reference synchronises the device stream on every start/stop (``internlm/utils/megatron_timers.py
:21-37``) which serialises host and device at B200 speeds; these timers record CUDA events on the current stream and
resolve elapsed time lazily, so timing a region costs two event records and no sync."""
from __future__ import annotations

import time
from collections import deque

import torch


class _Timer:
    def __init__(self, name):
        self.name_ = name
        self.elapsed_ = 0.0
        self.started_ = False
        self.start_time = 0.0
        self._pending = []  # (start_event, end_event)
        self._start_event = None
        self.history = deque(maxlen=10)

    def has_history(self):
        return len(self.history) > 0

    def start(self, reset_all=True):
        assert not self.started_, f"timer {self.name_} has already been started"
        if torch.cuda.is_available():
            self._start_event = torch.cuda.Event(enable_timing=True)
            self._start_event.record()
        self.start_time = time.time()
        self.started_ = True

    def stop(self):
        assert self.started_, f"timer {self.name_} is not started"
        if torch.cuda.is_available():
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self._pending.append((self._start_event, end))
        else:
            self.elapsed_ += time.time() - self.start_time
        self.started_ = False

    def _resolve(self):
        still = []
        for s, e in self._pending:
            if e.query():
                self.elapsed_ += s.elapsed_time(e) / 1000.0
            else:
                still.append((s, e))
        self._pending = still

    def reset(self):
        self.elapsed_ = 0.0
        self.started_ = False
        self._pending = []

    def elapsed(self, reset=True, sync=True):
        started = self.started_
        if started:
            self.stop()
        if self._pending and sync:
            self._pending[-1][1].synchronize()
        self._resolve()
        elapsed = self.elapsed_
        if reset:
            self.history.append(elapsed)
            self.reset()
        if started:
            self.start()
        return elapsed


class Timers:
    def __init__(self):
        self.timers = {}
        self.hist = {}
        self.names = []
        self.times = []

    def __call__(self, name):
        if name not in self.timers:
            self.timers[name] = _Timer(name)
        return self.timers[name]

    def store_last_timers(self):
        self.names, self.times = [], []
        for name, t in self.timers.items():
            self.names.append(name)
            self.times.append(t.elapsed(reset=False))

    def write(self, names, writer, iteration, normalizer=1.0, reset=False):
        assert normalizer > 0.0
        for name in names:
            if name in self.timers:
                writer.add_scalar(f"time/{name}-time", self.timers[name].elapsed(reset=reset) / normalizer, iteration)

    def log(self, names, logger, normalizer=1.0, reset=True):
        assert normalizer > 0.0
        string = ""
        for name in names:
            if name in self.timers:
                string += " | {}: {:.2f}".format(name, self.timers[name].elapsed(reset=reset) * 1000.0 / normalizer)
        if string:
            logger.info("time (ms)" + string)
        return string

    def debug(self, names, logger, normalizer=1.0, reset=True):
        return self.log(names, logger, normalizer, reset)

    def reset(self):
        for t in self.timers.values():
            t.reset()


megatron_timer = Timers()
