"""Named stopwatches for the log lines (same behaviour as kevlar/timer.py:13-39)."""
import time

_DEFAULT = ''


class Timer(object):
    """`start(key)` / `stop(key)` / `probe(key)`; the unnamed stopwatch is the key ''."""

    def __init__(self):
        self._began = {}
        self._ended = {}

    def _require(self, key):
        key = _DEFAULT if key is None else key
        if key not in self._began:
            raise ValueError('No timer started for "{}"'.format(key))
        return key

    def start(self, key=None):
        key = _DEFAULT if key is None else key
        if key in self._began:
            raise ValueError('Timer already started for "{}"'.format(key))
        self._began[key] = time.time()

    def stop(self, key=None):
        key = self._require(key)
        self._ended[key] = now = time.time()
        return now - self._began[key]

    def probe(self, key=None):
        return time.time() - self._began[self._require(key)]
