"""Features container: data matrix + timestamps + properties

Counterpart of shennong/features.py (validation rules :298-348, concatenation
:350-437).  Data produced by the GPU engine arrives here as a fresh numpy
array (one device -> host copy per batch).
"""

import copy

import numpy as np

from shennong_b200.logger import get_logger
from shennong_b200.utils import dict_equal


class Features:
    """`data` [nframes, ndims], `times` [nframes] or [nframes, 2], and a
    `properties` dict describing how the features were computed"""
    def __init__(self, data, times, properties=None, validate=True):
        self._data = data
        self._times = times
        self._properties = {} if properties is None else properties
        if validate is True:
            self.validate()

    @classmethod
    def _deferred(cls, data, times, properties):
        """Features whose `times` and `properties` are zero-argument
        callables evaluated at first access: the batch entry points wrap the
        rows of tens of thousands of utterances without building their
        timestamps and property dicts up front"""
        return cls(data, times, properties, validate=False)

    data = property(lambda self: self._data, doc='the features matrix')

    @property
    def times(self):
        """frames timestamps"""
        if callable(self._times):
            self._times = self._times()
        return self._times

    @property
    def properties(self):
        """metadata of the features"""
        if callable(self._properties):
            self._properties = self._properties()
        return self._properties

    dtype = property(lambda self: self.data.dtype)
    shape = property(lambda self: self.data.shape)
    ndims = property(lambda self: self.shape[1])
    nframes = property(lambda self: self.shape[0])

    def _to_dict(self, with_properties=True):
        out = {'data': self.data, 'times': self.times}
        if with_properties:
            out['properties'] = self.properties
        return out

    @staticmethod
    def _from_dict(features, validate=True):
        missing = {'data', 'times'} - set(features.keys())
        if missing:
            raise ValueError(
                'cannot read features from dict, missing keys: {}'
                .format(', '.join(missing)))
        return Features(
            features['data'], features['times'],
            properties=features.get('properties', {}), validate=validate)

    def __eq__(self, other):
        if self is other:
            return True
        return (self.shape == other.shape and self.dtype == other.dtype
                and dict_equal(self.properties, other.properties)
                and np.array_equal(self.times, other.times)
                and np.array_equal(self.data, other.data))

    def is_close(self, other, rtol=1e-5, atol=1e-8):
        """True if the features are equal up to numerical tolerance"""
        if self is other:
            return True
        return bool(
            self.shape == other.shape
            and dict_equal(self.properties, other.properties)
            and np.array_equal(self.times, other.times)
            and np.allclose(self.data, other.data, atol=atol, rtol=rtol))

    def copy(self, dtype=None, subsample=None):
        """Deep copy, optionally converted to `dtype` and/or keeping one
        frame every `subsample`"""
        if subsample is None:
            subsample = 1
        elif not isinstance(subsample, int) or subsample <= 0:
            raise ValueError(
                f'subsample must be a strictly positive integer, '
                f'it is: {subsample}')
        data, times = self.data[::subsample], self.times[::subsample]
        if dtype:
            data, times = data.astype(dtype), times.astype(dtype)
        else:
            data, times = data.copy(), times.copy()
        return Features(
            data, times, properties=copy.deepcopy(self.properties),
            validate=False)

    def is_valid(self):
        try:
            self.validate()
        except ValueError:
            return False
        return True

    def validate(self):
        """Raises ValueError if the features are inconsistent"""
        errors = []
        if not isinstance(self.data, np.ndarray):
            errors.append('data must be a numpy array')
        if not isinstance(self.times, np.ndarray):
            errors.append('times must be a numpy array')
        if not isinstance(self.properties, dict):
            errors.append('properties must be a dictionnary')
        if errors:
            raise ValueError(
                'invalid features data types: {}'.format(', '.join(errors)))
        if self.data.ndim != 2:
            errors.append(
                'data dimension must be 2 but is {}'.format(self.data.ndim))
        if self.times.ndim > 2:
            errors.append('times dimension must be 1 or 2 but is {}'.format(
                self.times.ndim))
        if self.times.ndim == 2 and self.times.shape[1] != 2:
            errors.append('times shape[1] must be 2, it is {}'.format(
                self.times.shape[1]))
        if self.data.shape[0] != self.times.shape[0]:
            errors.append(
                'mismatch in number of frames: {} for data but {} for times'
                .format(self.data.shape[0], self.times.shape[0]))
        if errors:
            raise ValueError(
                'invalid features dimensions: {}'.format(', '.join(errors)))
        order = (np.argsort(self.times, kind='stable') if self.times.ndim == 1
                 else np.lexsort(self.times.T))
        if not np.array_equal(order, np.arange(self.nframes)):
            raise ValueError('times is not sorted in increasing order')
        if not np.all(np.isfinite(self.data)):
            raise ValueError(
                'data contains non-finite numbers (nan of infinity)')

    def concatenate(self, other, tolerance=0,
                    log=get_logger('features', 'info')):
        """Column-wise concatenation with `other`

        The longest features are trimmed when the frame counts differ by at
        most `tolerance` frames (used to paste pitch, pipeline.py:639-641).
        Raises ValueError on larger differences or on different timestamps.
        """
        diff = abs(self.nframes - other.nframes)
        data1, data2 = self.data, other.data
        times1, times2 = self.times, other.times
        if diff:
            if not tolerance:
                raise ValueError('features have a different number of frames')
            if diff > tolerance:
                raise ValueError(
                    'features differs number of frames, and greater than '
                    'tolerance: |{} - {}| > {}'.format(
                        self.nframes, other.nframes, tolerance))
            log.warning(
                'features differs in number of frames, but within tolerance '
                '(|%s - %s| <= %s), trim the longest one',
                self.nframes, other.nframes, tolerance)
            if self.nframes > other.nframes:
                data1, times1 = data1[:-diff], times1[:-diff]
            else:
                data2, times2 = data2[:-diff], times2[:-diff]
        if not np.allclose(times1, times2):
            raise ValueError('times are not equal')
        properties = copy.deepcopy(self.properties)
        theirs = copy.deepcopy(other.properties)
        properties.update(
            {k: v for k, v in theirs.items() if k != 'pipeline'})
        properties.setdefault('pipeline', [])
        for entry in theirs.get('pipeline', []):
            entry['columns'] = [c + self.ndims for c in entry['columns']]
            properties['pipeline'].append(entry)
        return Features(
            np.hstack((data1, data2)), times1, properties=properties)
