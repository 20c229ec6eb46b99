"""Logging in the reference's line format `[I::module::func] msg`
(xcltk/utils/xlog.py:7-89; basefc logs to stderr, rdr/fc/main.py:84)."""

import logging

_LEVEL = {logging.DEBUG: "D", logging.INFO: "I", logging.WARNING: "W",
          logging.ERROR: "E", logging.CRITICAL: "C"}


class XFormatter(logging.Formatter):
    def __init__(self, datefmt=None):
        super().__init__(fmt=None, datefmt=datefmt)

    def format(self, record):
        record.message = record.getMessage()
        head = "[" + _LEVEL.get(record.levelno, "U")
        if record.module:
            head += "::%s" % record.module
        if record.funcName:
            head += "::%s" % record.funcName
        if self.datefmt:
            head += "::" + self.formatTime(record, self.datefmt)
        s = head + "] " + (str(record.message) if record.message else "")
        if record.exc_info and not record.exc_text:
            record.exc_text = self.formatException(record.exc_info)
        if record.exc_text:
            s = s.rstrip("\n") + "\n" + record.exc_text
        if record.stack_info:
            s = s.rstrip("\n") + "\n" + self.formatStack(record.stack_info)
        return s


def init_logging(log_file=None, stream=None, fh_level=logging.DEBUG,
                 fh_datefmt="%Y-%m-%d %H:%M:%S", ch_level=logging.INFO, ch_datefmt=None):
    if log_file is None and stream is None:
        raise ValueError("at least one of 'log_file' and 'stream' should not be None.")
    handlers = []
    if log_file:
        fh = logging.FileHandler(log_file, mode="w")
        fh.setLevel(fh_level)
        fh.setFormatter(XFormatter(datefmt=fh_datefmt))
        handlers.append(fh)
    if stream:
        ch = logging.StreamHandler(stream=stream)
        ch.setLevel(ch_level)
        ch.setFormatter(XFormatter(datefmt=ch_datefmt))
        handlers.append(ch)
    logging.basicConfig(level=logging.DEBUG, handlers=handlers)
