"""Spectrogram (log power spectrum) features

Counterpart of shennong/processor/spectrogram.py: column 0 holds the frame
log-energy, ``ndims = padded_window_size / 2 + 1``.
"""

from shennong_b200 import _lib
from shennong_b200.base import Option, f32_py
from shennong_b200.features import Features
from shennong_b200.processor.base import FramesProcessor, stream_features


class SpectrogramProcessor(FramesProcessor):
    """Spectogram"""
    energy_floor = Option(
        'Floor on energy (absolute, not relative) in spectrogram '
        'computation', **f32_py())
    raw_energy = Option(
        'If true, compute energy before preemphasis and windowing',
        store=bool)

    def __init__(self, sample_rate=16000, frame_shift=0.01,
                 frame_length=0.025, dither=1.0, preemph_coeff=0.97,
                 remove_dc_offset=True, window_type='povey',
                 round_to_power_of_two=True, blackman_coeff=0.42,
                 snip_edges=True, energy_floor=0.0, raw_energy=True):
        super().__init__(
            sample_rate=sample_rate, frame_shift=frame_shift,
            frame_length=frame_length, dither=dither,
            preemph_coeff=preemph_coeff, remove_dc_offset=remove_dc_offset,
            window_type=window_type,
            round_to_power_of_two=round_to_power_of_two,
            blackman_coeff=blackman_coeff, snip_edges=snip_edges)
        self.energy_floor = energy_floor
        self.raw_energy = raw_energy

    @property
    def name(self):
        return 'spectrogram'

    @property
    def ndims(self):
        padded = _lib.lib().snb_padded_window_size(
            _lib.ref(self._frame_opts()))
        return int(padded / 2 + 1)

    def _feat_opts(self):
        return _lib.FeatOpts(
            kind=_lib.FEATURE_KINDS['spectrogram'],
            energy_floor=self.energy_floor, raw_energy=int(self.raw_energy))

    def _wrap(self, data):
        return Features(
            data, self.times(data.shape[0]),
            properties=self.get_properties())

    def process(self, signal):
        """Spectrogram of a mono `signal` (ValueError on channel or sample
        rate mismatch)"""
        return self._wrap(self._extract([signal])[0])

    def _process_batch(self, audios):
        return [self._wrap(d) for d in self._extract(audios)]

    def _process_stream(self, utts, njobs):
        return stream_features(self, utts, njobs, with_warp=False)
