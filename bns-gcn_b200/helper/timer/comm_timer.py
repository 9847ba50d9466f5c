"""``CommTimer`` (helper/timer/comm_timer.py:6-33 of the reference): named intervals; ``tot_time`` is what the log
line reports as ``Comm(s)``.  The reference brackets each exchange with the host clock and no device synchronisation;
for an exchange that lives on a CUDA stream that would time the launch only, so an interval opened with a ``stream``
is a pair of CUDA events on it (device time of the exchange, hidden behind compute or not).  Without a stream the
host clock is used, as in the reference."""
import time
from contextlib import contextmanager

import torch


class CommTimer(object):

    def __init__(self):
        self._seconds = {}          # interval name -> zero-argument callable giving its length in seconds

    @contextmanager
    def timer(self, name, stream=None):
        if name in self._seconds:
            raise Exception(name + " already exists")          # comm_timer.py:14-15
        if stream is None:
            begin = time.time()
            yield
            length = time.time() - begin
            self._seconds[name] = lambda: length
            return
        first, last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        first.record(stream)
        yield
        last.record(stream)

        def device_seconds():
            last.synchronize()       # call after the epoch's device work was waited for
            return first.elapsed_time(last) * 1e-3
        self._seconds[name] = device_seconds

    def tot_time(self):
        return sum((length() for length in self._seconds.values()), 0.0)

    def print_time(self, rank=0):
        for name, length in self._seconds.items():
            print(f'(rank {rank}) Communication time of {name}: {length()} seconds.')

    def clear(self):
        self._seconds = {}
