"""Window functions (counterpart of shennong/window.py)

The table is produced by ``snb_window_function`` (double math, float32
storage, as Kaldi's FeatureWindowFunction) -- the same host routine that fills
the tables of the CUDA plans.
"""

import ctypes

import numpy as np

from shennong_b200 import _lib


def types():
    """The supported window functions"""
    return sorted(['povey', 'hanning', 'hamming', 'rectangular', 'blackman'])


def window(length, type='povey', blackman_coeff=0.42):
    """Window of `type` with `length` samples

    Raises ValueError if `length` <= 0 or `type` is unknown.  Length 1 and
    length 2 (povey, blackman, hanning) are special-cased to ones like the
    reference does (shennong/window.py:97-105).
    """
    if int(length) <= 0:
        raise ValueError(
            'length must be strictly positive but is {}'.format(length))
    if type not in types():
        raise ValueError(
            'type must be in {} but is {}'.format(types(), type))
    if length == 1:
        return np.ones((1,))
    if length == 2 and type in ('povey', 'blackman', 'hanning'):
        return np.ones((2,))
    opts = _lib.make_frame_opts(
        1000, 10.0, length, 0.0, 0.0, False, type, False, blackman_coeff,
        True)
    out = np.zeros(int(length), dtype=np.float32)
    _lib.check(_lib.lib().snb_window_function(
        _lib.ref(opts), _lib.np_ptr(out), ctypes.c_int32(int(length))))
    return out
