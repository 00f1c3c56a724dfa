"""Base class of all processors: sklearn-like parameter handling

Mirrors shennong/base.py:9-150 (``get_params`` / ``set_params`` driven by the
``__init__`` signature) and adds :class:`Option`, a declarative descriptor the
processors use to expose their parameters: each option is a read/write
property with the cast the reference applies (Kaldi stores floats as float32)
and a docstring the pipeline turns into YAML comments.
"""

import abc
import collections
import inspect

import numpy as np

from shennong_b200.logger import get_logger


class Option:
    """A processor parameter: validated at set time, typed at get time

    Parameters
    ----------
    doc : str
        Documentation of the option (used for commented YAML configurations)
    store : callable, optional
        Conversion applied when the value is set (e.g. ``np.float32``, what a
        Kaldi options struct does to a python float)
    load : callable, optional
        Conversion applied when the value is read (e.g. ``np.float32`` for
        the getters that wrap the value, ``float`` for those which don't)
    check : callable, optional
        ``check(processor, value)`` raising ValueError on invalid values
    """
    def __init__(self, doc, store=None, load=None, check=None):
        self.__doc__ = doc
        self._store = store
        self._load = load
        self._check = check
        self.name = None

    def __set_name__(self, owner, name):
        self.name = name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        value = obj.__dict__['_options'][self.name]
        return self._load(value) if self._load else value

    def __set__(self, obj, value):
        if self._check:
            value = self._check(obj, value) or value
        if self._store:
            value = self._store(value)
        obj.__dict__.setdefault('_options', {})[self.name] = value


def ms_store(seconds):
    """seconds -> Kaldi's float32 milliseconds (processor/base.py:162-172)"""
    return np.float32(seconds * 1000.0)


def ms_load_f32(ms):
    """float32 milliseconds -> np.float32 seconds (processor/base.py:159)"""
    return np.float32(float(ms) / 1000.0)


def ms_load(ms):
    """float32 milliseconds -> python float seconds (frames.py:70)"""
    return float(ms) / 1000.0


def f32_f32():
    return dict(store=np.float32, load=np.float32)


def f32_py():
    """Stored as a Kaldi float32 but read back as a python float"""
    return dict(store=np.float32, load=float)


class BaseProcessor:
    """Base class of all the processors"""
    def __init__(self):
        self._logger = get_logger(self.name, level='info')

    def __repr__(self):
        return self.__class__.__name__

    @property
    @abc.abstractmethod
    def name(self):
        """Processor name"""

    @property
    def log(self):
        """Processor logger"""
        return self._logger

    def set_logger(self, level,
                   formatter='%(levelname)s - %(name)s - %(message)s'):
        """Changes level and/or format of the processor's logger"""
        self._logger = get_logger(self.name, level=level, formatter=formatter)

    @classmethod
    def _get_param_names(cls):
        """Sorted names of the arguments of ``__init__``"""
        init = cls.__init__
        if init is object.__init__:  # pragma: nocover
            return []
        names = []
        for param in inspect.signature(init).parameters.values():
            if param.name == 'self' or param.kind == param.VAR_KEYWORD:
                continue
            if param.kind == param.VAR_POSITIONAL:
                raise RuntimeError(
                    f'processors must specify their parameters in the '
                    f'signature of __init__ (no varargs): {cls}')
            names.append(param.name)
        return sorted(names)

    def get_params(self, deep=True):
        """Parameters of the processor as a dict {name: value}"""
        params = {}
        for key in self._get_param_names():
            value = getattr(self, key, None)
            if deep and hasattr(value, 'get_params'):
                params.update(
                    (key + '__' + k, v) for k, v in value.get_params().items())
            params[key] = value
        return params

    def set_params(self, **params):
        """Sets parameters of the processor, returns self

        Raises ValueError for unknown parameters.
        """
        if not params:
            return self
        valid = self.get_params(deep=True)
        nested = collections.defaultdict(dict)
        for key, value in params.items():
            key, delim, sub_key = key.partition('__')
            if key not in valid:
                raise ValueError(
                    f'invalid parameter {key} for processor {self}, '
                    f'check the list of available parameters '
                    f'with `processor.get_params().keys()`.')
            if delim:
                nested[key][sub_key] = value
            else:
                try:
                    setattr(self, key, value)
                except AttributeError:
                    raise ValueError(
                        f'cannot set attribute {key} for {self}') from None
                valid[key] = value
        for key, sub_params in nested.items():
            valid[key].set_params(**sub_params)
        return self
