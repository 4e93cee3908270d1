"""`_elf._logging`: the logger surface the reference's Python layer uses (src_cpp/elf/logging/Pybind.cc,
Levels.cc, IndexedLoggerFactory.cc -- spdlog there), on the standard `logging` module: LoggerLevel with
from_str, Logger objects with trace..critical, the registry functions, and getIndexedLogger /
IndexedLoggerFactory, which number the loggers they create."""
import enum
import itertools
import logging as _pylog
import sys
import threading


class LoggerLevel(enum.IntEnum):
    trace = 0
    debug = 1
    info = 2
    warn = 3
    err = 4
    critical = 5
    off = 6
    invalid = 127

    @staticmethod
    def from_str(s):
        names = ["trace", "debug", "info", "warning", "error", "critical", "off"]  # spdlog::level::level_names
        return LoggerLevel(names.index(s)) if s in names else LoggerLevel.invalid


_TO_PY = {0: 5, 1: _pylog.DEBUG, 2: _pylog.INFO, 3: _pylog.WARNING, 4: _pylog.ERROR, 5: _pylog.CRITICAL, 6: 1000}
_registry = {}
_lock = threading.Lock()
_global_level = [LoggerLevel.info]
_pattern = ["[%(asctime)s] [%(name)s] [%(levelname)s] %(message)s"]


class Logger:
    def __init__(self, name, stream=None):
        self._name = name
        self._log = _pylog.getLogger("elf." + name)
        self._log.propagate = False
        if not self._log.handlers:
            h = _pylog.StreamHandler(stream or sys.stdout)
            h.setFormatter(_pylog.Formatter(_pattern[0]))
            self._log.addHandler(h)
        self._level = _global_level[0]
        self._log.setLevel(_TO_PY[int(self._level)])

    def _emit(self, lvl, msg):
        if self.should_log(lvl):
            self._log.log(max(_TO_PY[int(lvl)], 1), msg)

    def trace(self, msg): self._emit(LoggerLevel.trace, msg)
    def debug(self, msg): self._emit(LoggerLevel.debug, msg)
    def info(self, msg): self._emit(LoggerLevel.info, msg)
    def warn(self, msg): self._emit(LoggerLevel.warn, msg)
    def error(self, msg): self._emit(LoggerLevel.err, msg)
    def critical(self, msg): self._emit(LoggerLevel.critical, msg)

    def flush(self):
        for h in self._log.handlers:
            h.flush()

    def flush_on(self, level):
        pass

    def level(self):
        return self._level

    def name(self):
        return self._name

    def set_formatter(self, f):
        pass

    def set_level(self, level):
        self._level = LoggerLevel(int(level))
        self._log.setLevel(_TO_PY[int(self._level)])

    def should_log(self, level):
        return int(level) >= int(self._level) and int(self._level) != int(LoggerLevel.off)


def _make(name, stream=None):
    with _lock:
        if name in _registry:
            raise RuntimeError(f"logger with name '{name}' already exists")
        lg = _registry[name] = Logger(name, stream)
        return lg


def get(name):
    return _registry.get(name)


def drop(name):
    _registry.pop(name, None)


def drop_all():
    _registry.clear()


def set_level(level):
    _global_level[0] = LoggerLevel(int(level))
    for lg in _registry.values():
        lg.set_level(level)


def set_pattern(pattern):
    pass  # spdlog pattern syntax; the stdlib format above stays


def stdout_logger_mt(name): return _make(name, sys.stdout)
def stderr_logger_mt(name): return _make(name, sys.stderr)
def stdout_color_mt(name): return _make(name, sys.stdout)
def stderr_color_mt(name): return _make(name, sys.stderr)
def daily_logger_mt(name, filename, hour=0, minute=0): return _make(name, open(filename, "a"))
def rotating_logger_mt(name, filename, max_size, max_files): return _make(name, open(filename, "a"))


class IndexedLoggerFactory:
    """IndexedLoggerFactory.h:29-64: makeLogger(prefix, suffix) -> creator(prefix + counter + suffix)"""

    def __init__(self, creator, init_index=0):
        self._creator = creator
        self._counter = itertools.count(init_index)

    def makeLogger(self, prefix, suffix):
        return self._creator(prefix + str(next(self._counter)) + suffix)


_default_factory = IndexedLoggerFactory(stderr_color_mt)


def getIndexedLogger(prefix, suffix):
    return _default_factory.makeLogger(prefix, suffix)
