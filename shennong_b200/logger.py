"""Logging helpers (same surface as shennong/logger.py:7-84)"""

import logging
import sys

_LEVELS = {'debug': logging.DEBUG, 'info': logging.INFO,
           'warning': logging.WARNING, 'error': logging.ERROR}


def null_logger(name='null'):
    """A logger that drops every message"""
    log = logging.getLogger(name)
    log.handlers = [logging.NullHandler()]
    return log


def get_logger(name, level,
               formatter='%(levelname)s - %(name)s - %(message)s'):
    """A logger writing to stderr at `level`

    `level` must be 'debug', 'info', 'warning' or 'error', else a ValueError
    is raised.
    """
    if level not in _LEVELS:
        raise ValueError(
            'invalid logging level "{}", must be in {}'.format(
                level, ', '.join(_LEVELS.keys())))
    handler = logging.StreamHandler(sys.stderr)
    handler.setFormatter(logging.Formatter(formatter))
    log = logging.getLogger(name)
    log.handlers = [handler]
    log.setLevel(_LEVELS[level])
    return log
