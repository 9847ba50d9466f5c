"""Thread-local ``comm_timer`` singleton (helper/timer/timer.py:3 of the reference)."""
import threading

from .comm_timer import CommTimer

_tls = threading.local()


class _TimerProxy:
    def _get(self):
        t = getattr(_tls, "t", None)
        if t is None:
            t = _tls.t = CommTimer()
        return t

    def __getattr__(self, item):
        return getattr(self._get(), item)


comm_timer = _TimerProxy()
