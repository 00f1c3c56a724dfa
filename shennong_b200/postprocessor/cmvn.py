"""Cepstral mean and variance normalisation
(counterpart of shennong/postprocessor/cmvn.py)

Statistics are float64 ``[2, dim+1]`` (sum, sum of squares, count) accumulated
on the device by ``snb_cmvn_accumulate`` (double sums of float products, fixed
order => reproducible); normalisation is ``snb_cmvn_apply``.
"""

import copy

import numpy as np

from shennong_b200 import engine
from shennong_b200.base import Option
from shennong_b200.features import Features
from shennong_b200.features_collection import FeaturesCollection
from shennong_b200.postprocessor.base import FeaturesPostProcessor


class CmvnPostProcessor(FeaturesPostProcessor):
    """CMVN statistics accumulation and application

    `dim` is the features dimension (strictly positive integer), `stats`
    optional pre-accumulated statistics shaped [2, dim+1].
    """
    def __init__(self, dim, stats=None):
        super().__init__()
        if not isinstance(dim, int) or dim <= 0:
            raise ValueError(
                'dimension must be a strictly positive integer, it is {}'
                .format(dim))
        self._dim = dim
        self._stats = np.zeros((2, dim + 1), dtype=np.float64)
        if stats is not None:
            stats = np.asarray(stats)
            if stats.shape != (2, dim + 1):
                raise ValueError(
                    'stats must be an array of shape {}, but is shaped as {}'
                    .format((2, dim + 1), stats.shape))
            self._stats = stats.astype(np.float64)

    @property
    def name(self):
        return 'cmvn'

    @property
    def dim(self):
        """The dimension of features on which to compute CMVN"""
        return self._dim

    @property
    def stats(self):
        """Accumulated statistics, [2, dim+1]: row 0 sums (and the count in
        the last column), row 1 sums of squares"""
        return self._stats

    @property
    def count(self):
        """The weighted count of accumulated frames"""
        return self.stats[0, -1]

    @property
    def ndims(self):
        return self.dim

    def get_properties(self, features):
        properties = super().get_properties(features)
        properties[self.name]['stats'] = self.stats
        return properties

    def accumulate(self, features, weights=None):
        """Accumulates the statistics of `features`, optionally weighted

        ValueError if `weights` is not 1d or its length differs from the
        number of frames.
        """
        w = None
        if weights is not None:
            if weights.ndim != 1:
                raise ValueError(
                    'weights must have a single dimension but have {}'
                    .format(weights.ndim))
            if weights.shape[0] != features.nframes:
                raise ValueError(
                    'there is {} weights but {} feature frames, must be equal'
                    .format(weights.shape[0], features.nframes))
            w = engine.from_host(weights, np.float32)
        if features.ndims != self.dim:
            raise ValueError(
                'features dimension is {} but cmvn dimension is {}'.format(
                    features.ndims, self.dim))
        x = engine.from_host(features.data, np.float32)
        layout = engine.RowLayout([0, x.shape[0]])
        stats = engine.to_host(engine.cmvn_accumulate(x, layout, w))[0]
        # new array: properties of previously returned features keep theirs
        self._stats = self._stats + stats

    def add_stats(self, stats):
        """Adds pre-computed statistics (used by the batched pipeline)"""
        self._stats = self._stats + np.asarray(stats, dtype=np.float64)

    def _effective_stats(self, skip_dims, ndims):
        if not skip_dims:
            return self._stats
        dmin, dmax = min(skip_dims), max(skip_dims)
        if dmin < 0 or dmax >= ndims:
            raise ValueError(
                'skipped dimensions must be in [0, {}[ but are in [{}, {}['
                .format(ndims, dmin, dmax))
        # FakeStatsForSomeDims: zero mean, unit variance on a copy
        stats = self._stats.copy()
        for d in skip_dims:
            stats[0, d] = 0.0
            stats[1, d] = stats[0, -1]
        return stats

    def process(self, features, norm_vars=True, skip_dims=None, reverse=False):
        """Applies the accumulated statistics to `features`

        ValueError when fewer than one frame has been accumulated.
        """
        if self.count < 1.0:
            raise ValueError(
                'insufficient accumulation of stats for CMVN, '
                'must be >= 1.0 but is {}'.format(self.count))
        stats = self._effective_stats(skip_dims, features.ndims)
        x = engine.from_host(features.data, np.float32)
        layout = engine.RowLayout([0, x.shape[0]])
        norm = engine.cmvn_norm(
            engine.from_host(stats[None], np.float64), norm_vars, reverse)
        out = engine.cmvn_apply(x, layout, norm)
        return Features(
            engine.to_host(out), features.times,
            properties=self.get_properties(features))


def apply_cmvn(feats_collection, by_collection=True, norm_vars=True,
               weights=None, skip_dims=None):
    """CMVN of a whole collection, globally or per item
    (shennong/postprocessor/cmvn.py:285-379)"""
    dims = set(f.ndims for f in feats_collection.values())
    if len(dims) != 1:
        raise ValueError(
            'features in the collection must have consistent dimensions '
            'but dimensions are: {}'.format(sorted(dims)))
    dim = dims.pop()
    if weights is not None and weights.keys() != feats_collection.keys():
        raise ValueError('keys differ for weights and features collection')
    if skip_dims is not None:
        lo, hi = min(skip_dims), max(skip_dims)
        if lo < 0 or hi >= dim:
            raise ValueError(
                'out of bounds dimensions in skip_dims, must be in [0, {}] '
                'but are in [{}, {}]'.format(dim - 1, lo, hi))

    def weight(key):
        return weights[key] if weights is not None else None

    if by_collection:
        cmvn = CmvnPostProcessor(dim)
        for key, feats in feats_collection.items():
            cmvn.accumulate(feats, weights=weight(key))
        return FeaturesCollection(
            {k: cmvn.process(f, norm_vars=norm_vars, skip_dims=skip_dims)
             for k, f in feats_collection.items()})
    out = FeaturesCollection()
    for key, feats in feats_collection.items():
        cmvn = CmvnPostProcessor(feats.ndims)
        cmvn.accumulate(feats, weights=weight(key))
        out[key] = cmvn.process(
            feats, norm_vars=norm_vars, skip_dims=skip_dims)
    return out


class SlidingWindowCmvnPostProcessor(FeaturesPostProcessor):
    """Sliding-window mean (and variance) normalisation"""
    center = Option('Whether to center the window on the current frame',
                    store=bool)
    cmn_window = Option('Window size for average CMN computation', store=int)
    min_window = Option('Minimum CMN window used at start of decoding',
                        store=int)
    max_warnings = Option('Maximum warning to report per utterance',
                          store=int)
    normalize_variance = Option('Whether to normalize variance to one',
                                store=bool)

    def __init__(self, center=True, cmn_window=600, min_window=100,
                 max_warnings=5, normalize_variance=False):
        super().__init__()
        self.center = center
        self.cmn_window = cmn_window
        self.max_warnings = max_warnings
        self.min_window = min_window
        self.normalize_variance = normalize_variance

    @property
    def name(self):
        return 'sliding_window_cmvn'

    @property
    def ndims(self):
        raise ValueError('output dimension for sliding '
                         'window CMVN processor depends on input')

    def get_properties(self, features):
        properties = copy.deepcopy(features.properties)
        properties[self.name] = self.get_params()
        properties.setdefault('pipeline', []).append(
            {'name': self.name, 'columns': [0, features.ndims - 1]})
        return properties

    def process(self, features):
        """Sliding-window normalisation of `features`"""
        x = engine.from_host(features.data, np.float32)
        layout = engine.RowLayout([0, x.shape[0]])
        out = engine.sliding_window_cmn(
            x, layout, self.center, self.cmn_window, self.min_window,
            self.normalize_variance)
        return Features(
            engine.to_host(out), features.times,
            self.get_properties(features))
