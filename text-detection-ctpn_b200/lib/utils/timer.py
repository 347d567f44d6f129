"""Interval timer behind the attribute names callers of the reference's lib/utils/timer.py read: tic() / toc(average),
.total_time, .calls, .diff, .average_time, .start_time.  Intervals come from the monotonic performance counter."""
from time import perf_counter


class Timer:
    __slots__ = ("total_time", "calls", "start_time", "diff")

    def __init__(self):
        self.total_time, self.calls, self.start_time, self.diff = 0.0, 0, 0.0, 0.0

    @property
    def average_time(self):
        return self.total_time / self.calls if self.calls else 0.0

    def tic(self):
        self.start_time = perf_counter()

    def toc(self, average=True):
        now = perf_counter()
        self.diff = now - self.start_time
        self.calls += 1
        self.total_time += self.diff
        return self.average_time if average else self.diff
