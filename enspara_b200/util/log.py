"""``timed`` context manager (reference: enspara/util/log.py:5-10)."""
import time
from contextlib import contextmanager


@contextmanager
def timed(fmt, log_func):
    t0 = time.perf_counter()
    yield
    log_func(fmt % (time.perf_counter() - t0))
