"""Delta features (counterpart of shennong/postprocessor/delta.py)

``out[t, i*d:(i+1)*d] = sum_j scales_i[j] * x[clamp(t + j)]`` with Kaldi's
composed-filter scales, computed by ``snb_compute_deltas``.
"""

import copy

import numpy as np

from shennong_b200 import engine
from shennong_b200.base import Option
from shennong_b200.features import Features
from shennong_b200.postprocessor.base import FeaturesPostProcessor


def _check_window(_, value):
    if not 0 < value < 1000:
        raise ValueError(
            'window must be in [1, 999], it is {}'.format(value))


class DeltaPostProcessor(FeaturesPostProcessor):
    """Time derivatives of the features"""
    order = Option('Order of delta computation', store=int)
    window = Option(
        'Parameter controlling window for delta computation\n\n'
        'The actual window size for each delta order is 1 + 2 * `window`. '
        'The behavior at the edges is to replicate the first or last frame.',
        store=int, check=_check_window)

    def __init__(self, order=2, window=2):
        super().__init__()
        self.order = order
        self.window = window

    @property
    def name(self):
        return 'delta'

    @property
    def ndims(self):
        raise ValueError(
            'output dimension for delta processor depends on input')

    def get_properties(self, features):
        properties = copy.deepcopy(features.properties)
        properties[self.name] = {'order': self.order, 'window': self.window}
        properties.setdefault('pipeline', []).append({
            'name': self.name,
            'columns': [0, (self.order + 1) * features.ndims - 1]})
        return properties

    def process(self, features):
        """Features [nframes, ncols] -> [nframes, ncols * (order + 1)]"""
        x = engine.from_host(features.data, np.float32)
        layout = engine.RowLayout([0, x.shape[0]])
        out = engine.deltas(x, layout, self.order, self.window)
        return Features(
            engine.to_host(out), features.times,
            self.get_properties(features))
