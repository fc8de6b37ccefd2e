"""Wall-clock logging helper used by the apps (the reference has a context manager of the same
name and call convention, enspara/util/log.py)."""
import time


class timed:
    """``with timed("Wrote centers in %.2f sec.", logger.info): ...`` logs the elapsed seconds
    when the block finishes normally; the measurement stays available as ``.seconds``."""

    def __init__(self, fmt, log_func):
        self.fmt, self.log_func, self.seconds = fmt, log_func, None

    def __enter__(self):
        self._start = time.perf_counter()
        return self

    def __exit__(self, exc_type, exc, tb):
        self.seconds = time.perf_counter() - self._start
        if exc_type is None:
            self.log_func(self.fmt % self.seconds)
        return False
