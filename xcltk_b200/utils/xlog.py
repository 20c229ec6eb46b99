"""Log lines shaped like the reference's: `[I::module::func] msg`, with `::<time>` appended
to the bracket when a date format is set (xcltk/utils/xlog.py:7-89; basefc logs to stderr,
rdr/fc/main.py:84)."""

import logging

_LETTER = dict(zip((logging.DEBUG, logging.INFO, logging.WARNING, logging.ERROR, logging.CRITICAL), "DIWEC"))


class LineFormatter(logging.Formatter):
    """One bracketed prefix, then the message; tracebacks / stack text follow on their own lines."""

    def __init__(self, datefmt=None):
        logging.Formatter.__init__(self, None, datefmt)

    def _prefix(self, rec):
        parts = [_LETTER.get(rec.levelno, "U")]
        parts.extend(x for x in (rec.module, rec.funcName) if x)
        if self.datefmt:
            parts.append(self.formatTime(rec, self.datefmt))
        return "[%s] " % "::".join(parts)

    def format(self, rec):
        rec.message = rec.getMessage()
        text = self._prefix(rec) + (str(rec.message) if rec.message else "")
        if rec.exc_info and not rec.exc_text:
            rec.exc_text = self.formatException(rec.exc_info)
        for extra in (rec.exc_text, self.formatStack(rec.stack_info) if rec.stack_info else None):
            if extra:
                text = text + ("" if text.endswith("\n") else "\n") + extra
        return text


XFormatter = LineFormatter      # the reference's name for it


def _attach(handler, level, datefmt, to):
    handler.setLevel(level)
    handler.setFormatter(LineFormatter(datefmt))
    to.append(handler)


def init_logging(log_file=None, stream=None, fh_level=logging.DEBUG, fh_datefmt="%Y-%m-%d %H:%M:%S",
                 ch_level=logging.INFO, ch_datefmt=None):
    """File handler (timestamps) and / or stream handler (none), root logger at DEBUG."""
    if log_file is None and stream is None:
        raise ValueError("at least one of 'log_file' and 'stream' should not be None.")
    sinks = []
    if log_file:
        _attach(logging.FileHandler(log_file, mode="w"), fh_level, fh_datefmt, sinks)
    if stream:
        _attach(logging.StreamHandler(stream), ch_level, ch_datefmt, sinks)
    logging.basicConfig(level=logging.DEBUG, handlers=sinks)
