"""``setup_logger(name)`` (mirror of cosmogrb/utils/logging.py:262-282, console handler only: the
reference's rotating log files under ~/.cosmogrb/log are infrastructure outside the hot path)."""
import logging

from ..config import cosmogrb_config

_handler = None


def _console_handler():
    global _handler
    if _handler is None:
        _handler = logging.StreamHandler()
        _handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)-8s %(name)s | %(message)s"))
        _handler.setLevel(getattr(logging, str(cosmogrb_config["logging"]["console"]["level"]).upper(), logging.WARNING))
    return _handler


def setup_logger(name):
    log = logging.getLogger(name)
    log.setLevel(logging.DEBUG if cosmogrb_config["logging"]["debug"] else
                 getattr(logging, str(cosmogrb_config["logging"]["level"]).upper(), logging.INFO))
    if cosmogrb_config["logging"]["console"]["on"] and _console_handler() not in log.handlers:
        log.addHandler(_console_handler())
    log.propagate = False
    return log


def update_logging_level(level):
    _console_handler().setLevel(level)
