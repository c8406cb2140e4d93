"""CUDA-event phase timers of empbayes_fit.

The reference stamps wall-clock at three points inside the jitted objective by threading a token through
jax.pure_callback (src/lsqfitgp/_fit.py:41-77,410-442): after the GP is built and its covariance assembled
('gp&cov'), after the decomposition ('decomp'), after the likelihood and its derivatives ('likelihood'), and reports the
totals (:775-794).  Here the same three marks are CUDA events recorded on the evaluation's stream by the GP code
(`mark()` calls in _GP.py); nothing synchronises until `stop()`, which the fit calls after it has read the result back
anyway.
"""

import threading
import time

import torch

PHASES = ('gp&cov', 'decomp', 'likelihood')
_current = threading.local()


class PhaseTimer:
    """ accumulates device time per phase over many evaluations; evaluations may run concurrently on several host
    threads (one stream each: empbayes_fit's multistart batch), so the events of an evaluation live in thread-local
    storage and only the totals are shared """

    def __init__(self):
        self.totals = dict.fromkeys(PHASES, 0.0)   # seconds of device time
        self.wall = 0.0
        self._lock = threading.Lock()

    @staticmethod
    def _event():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def start(self):
        _current.timer = self
        _current.events = [(None, self._event())] if torch.cuda.is_available() else None
        _current.t0 = time.perf_counter()

    def mark(self, name):
        events = getattr(_current, 'events', None)
        if events is not None and events[-1][0] != name:
            events.append((name, self._event()))

    def stop(self):
        events = getattr(_current, 'events', None)
        t0 = getattr(_current, 't0', None)
        _current.timer = _current.events = _current.t0 = None
        if t0 is None:
            return
        wall = time.perf_counter() - t0
        parts = {}
        if events is not None:
            # whatever follows the last mark (value, gradient) is the likelihood phase
            events.append(('likelihood', self._event()))
            events[-1][1].synchronize()
            for (_, e0), (name, e1) in zip(events, events[1:]):
                parts[name] = parts.get(name, 0.0) + e0.elapsed_time(e1) * 1e-3
        with self._lock:
            self.wall += wall
            for k, v in parts.items():
                self.totals[k] += v


def mark(name):
    """ called by the GP code at the end of a phase; no-op outside a timed evaluation """
    t = getattr(_current, 'timer', None)
    if t is not None:
        t.mark(name)
