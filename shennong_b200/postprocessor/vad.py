"""Energy based voice activity detection
(counterpart of shennong/postprocessor/vad.py)"""

import numpy as np

from shennong_b200 import engine
from shennong_b200.base import Option, f32_f32
from shennong_b200.features import Features
from shennong_b200.postprocessor.base import FeaturesPostProcessor


def _non_negative(what):
    def check(_, value):
        if value < 0:
            raise ValueError(f'{what} must be >= 0, it is {value}')
    return check


def _check_proportion(_, value):
    if value <= 0 or value >= 1:
        raise ValueError(
            'proportion_threshold must be in ]0, 1[, it is {}'.format(value))


class VadPostProcessor(FeaturesPostProcessor):
    """Computes VAD on speech features"""
    energy_threshold = Option(
        'Constant term in energy threshold for MFCC0 for VAD\n\n'
        'See also :func:`energy_mean_scale`', **f32_f32())
    energy_mean_scale = Option(
        'Scale factor of the mean log-energy\n\n'
        'If this is set to `s`, to get the actual threshold we let `m` be '
        'the mean log-energy of the file, and use `s*m +` '
        ':func:`energy_threshold`. Must be greater or equal to 0.',
        check=_non_negative('Energy mean scale'), **f32_f32())
    frames_context = Option(
        'Number of frames of context on each side of central frame\n\n'
        'The size of the window for which energy is monitored is '
        '`2 * frames_context + 1`. Must be greater or equal to 0.',
        store=int, check=_non_negative('frames_context'))
    proportion_threshold = Option(
        'Proportion of frames beyond the energy threshold\n\n'
        'Parameter controlling the proportion of frames within the window '
        'that need to have more energy than the threshold. Must be in '
        ']0, 1[.', check=_check_proportion, **f32_f32())

    def __init__(self, energy_threshold=5.0, energy_mean_scale=0.5,
                 frames_context=0, proportion_threshold=0.6):
        super().__init__()
        self.energy_threshold = energy_threshold
        self.energy_mean_scale = energy_mean_scale
        self.frames_context = frames_context
        self.proportion_threshold = proportion_threshold

    @property
    def name(self):
        return 'vad'

    @property
    def ndims(self):
        return 1

    def process(self, features):
        """uint8 [nframes, 1]: 1 for voiced frames, 0 otherwise; the first
        column of `features` must be a log-energy"""
        x = engine.from_host(features.data, np.float32)
        layout = engine.RowLayout([0, x.shape[0]])
        out = engine.vad_energy(
            x, layout, self.energy_threshold, self.energy_mean_scale,
            self.frames_context, self.proportion_threshold)
        data = engine.to_host(out).astype(np.uint8)
        return Features(
            np.atleast_2d(data).T, features.times,
            properties=self.get_properties(features))
