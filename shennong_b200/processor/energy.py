"""Frame energy (counterpart of shennong/processor/energy.py)

Per frame ``compression(sum(frame ** 2))`` accumulated in float64, the frame
being extracted with the processor's options; ``raw_energy`` disables
pre-emphasis and windowing (energy.py:148-151).  Like the reference the signal
is NOT cast to int16 (energy.py:158): float audio keeps its [-1, 1] scale.
"""

import numpy as np

from shennong_b200 import _lib
from shennong_b200.base import Option
from shennong_b200.features import Features
from shennong_b200.processor.base import FramesProcessor

_COMPRESSIONS = ('off', 'log', 'sqrt')


def _check_compression(_, value):
    if value not in _COMPRESSIONS:
        raise ValueError(
            'compression must be in {}, it is {}'.format(
                ', '.join(_COMPRESSIONS), value))


class EnergyProcessor(FramesProcessor):
    """Energy of the frames of an audio signal"""
    raw_energy = Option(
        'If true, compute energy before preemphasis and windowing')
    compression = Option(
        "Type of energy compression\n\nMust be 'off' (disable compression), "
        "'log' (natural logarithm) or 'sqrt' (squared root).",
        check=_check_compression)

    def __init__(self, sample_rate=16000, frame_shift=0.01,
                 frame_length=0.025, dither=1.0, preemph_coeff=0.97,
                 remove_dc_offset=True, window_type='povey',
                 round_to_power_of_two=True, blackman_coeff=0.42,
                 snip_edges=True, raw_energy=True, compression='log'):
        super().__init__(
            sample_rate=sample_rate, frame_shift=frame_shift,
            frame_length=frame_length, dither=dither,
            preemph_coeff=preemph_coeff, remove_dc_offset=remove_dc_offset,
            window_type=window_type,
            round_to_power_of_two=round_to_power_of_two,
            blackman_coeff=blackman_coeff, snip_edges=snip_edges)
        self.compression = compression
        self.raw_energy = raw_energy

    @property
    def name(self):
        return 'energy'

    @property
    def ndims(self):
        return 1

    def _feat_opts(self):
        return _lib.FeatOpts(
            kind=_lib.FEATURE_KINDS['energy'],
            raw_energy=int(bool(self.raw_energy)),
            energy_compression=_lib.ENERGY_COMPRESSION[self.compression])

    def _output_float64(self):
        return True

    def _pcm(self, signal):
        data = signal.data
        if data.dtype == np.int16:
            return data, np.int16
        return np.asarray(data, dtype=np.float32), np.float32

    def _wrap(self, data):
        return Features(
            data, self.times(data.shape[0]), self.get_properties())

    def process(self, signal):
        """Energy of a mono `signal`, float64 [nframes, 1]"""
        return self._wrap(self._extract([signal])[0])

    def _process_batch(self, audios):
        return [self._wrap(d) for d in self._extract(audios)]
