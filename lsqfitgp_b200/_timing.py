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
    def __init__(self):
        self.totals = dict.fromkeys(PHASES, 0.0)   # seconds of device time
        self.wall = 0.0
        self._events = None

    def _event(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def start(self):
        self._events = [(None, self._event())] if torch.cuda.is_available() else None
        self._t0 = time.perf_counter()
        _current.timer = self

    def mark(self, name):
        if self._events is not None and self._events[-1][0] != name:
            self._events.append((name, self._event()))

    def stop(self):
        _current.timer = None
        self.wall += time.perf_counter() - self._t0
        if self._events is None:
            return
        # whatever follows the last mark (value, gradient) is the likelihood phase
        self._events.append(('likelihood', self._event()))
        self._events[-1][1].synchronize()
        for (_, e0), (name, e1) in zip(self._events, self._events[1:]):
            self.totals[name] += e0.elapsed_time(e1) * 1e-3
        self._events = None


def mark(name):
    """ called by the GP code at the end of a phase; no-op outside a timed evaluation """
    t = getattr(_current, 'timer', None)
    if t is not None:
        t.mark(name)
