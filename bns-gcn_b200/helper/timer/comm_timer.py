"""``CommTimer`` (helper/timer/comm_timer.py:6-33 of the reference): named intervals, ``tot_time`` is what
the log line reports as ``Comm(s)``.  The reference brackets each exchange with host wall-clock and no device
synchronisation; on a GPU-resident exchange that would time only the launch, so every interval here is a pair of
CUDA events on the communication stream (device time of the exchange itself, whether or not it was hidden
behind compute) -- the host clock is kept for CPU use."""
import time
from contextlib import contextmanager

import torch


class CommTimer(object):

    def __init__(self):
        super(CommTimer, self).__init__()
        self._time = {}
        self._events = {}

    @contextmanager
    def timer(self, name, stream=None):
        if name in self._time or name in self._events:
            raise Exception(name + " already exists")          # comm_timer.py:14-15
        if stream is not None:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            yield
            e.record(stream)
            self._events[name] = (s, e)
        else:
            t0 = time.time()
            yield
            self._time[name] = (t0, time.time())

    def tot_time(self):
        """Seconds.  Synchronises the recorded events (call it after the epoch's device work was waited for)."""
        tot = 0.0
        for (t0, t1) in self._time.values():
            tot += t1 - t0
        for (s, e) in self._events.values():
            e.synchronize()
            tot += s.elapsed_time(e) * 1e-3
        return tot

    def print_time(self, rank=0):
        for (k, (t0, t1)) in self._time.items():
            print(f'(rank {rank}) Communication time of {k}: {t1 - t0} seconds.')
        for (k, (s, e)) in self._events.items():
            e.synchronize()
            print(f'(rank {rank}) Communication time of {k}: {s.elapsed_time(e) * 1e-3} seconds.')

    def clear(self):
        self._time = {}
        self._events = {}
