"""Mel filterbank features (counterpart of shennong/processor/filterbank.py)"""

from shennong_b200 import _lib
from shennong_b200.base import Option, f32_py
from shennong_b200.processor.base import MelFeaturesProcessor


class FilterbankProcessor(MelFeaturesProcessor):
    """Mel-filterbank features"""
    use_energy = Option(
        'Add an extra dimension with energy to the filterbank output',
        store=bool)
    energy_floor = Option(
        'Floor on energy (absolute, not relative) in filterbanks',
        **f32_py())
    raw_energy = Option(
        'If true, compute energy before preemphasis and windowing',
        store=bool)
    htk_compat = Option(
        'If True, get closer to HTK filterbank features.\n\n'
        'Put energy last.\n\nWarning: Not sufficient to get HTK compatible '
        'features (need to change other parameters)', store=bool)
    use_log_fbank = Option(
        'If true, produce log-filterbank, else produce linear', store=bool)
    use_power = Option('If true, use power, else use magnitude', store=bool)

    def __init__(self, sample_rate=16000, frame_shift=0.01,
                 frame_length=0.025, dither=1.0, preemph_coeff=0.97,
                 remove_dc_offset=True, window_type='povey',
                 round_to_power_of_two=True, blackman_coeff=0.42,
                 snip_edges=True, num_bins=23, low_freq=20,
                 high_freq=0, vtln_low=100, vtln_high=-500,
                 use_energy=False, energy_floor=0.0, raw_energy=True,
                 htk_compat=False, use_log_fbank=True, use_power=True):
        super().__init__(
            sample_rate=sample_rate, frame_shift=frame_shift,
            frame_length=frame_length, dither=dither,
            preemph_coeff=preemph_coeff, remove_dc_offset=remove_dc_offset,
            window_type=window_type,
            round_to_power_of_two=round_to_power_of_two,
            blackman_coeff=blackman_coeff, snip_edges=snip_edges,
            num_bins=num_bins, low_freq=low_freq, high_freq=high_freq,
            vtln_low=vtln_low, vtln_high=vtln_high)
        self.use_energy = use_energy
        self.energy_floor = energy_floor
        self.raw_energy = raw_energy
        self.htk_compat = htk_compat
        self.use_log_fbank = use_log_fbank
        self.use_power = use_power

    @property
    def name(self):
        return 'filterbank'

    @property
    def ndims(self):
        return self.num_bins + 1 if self.use_energy else self.num_bins

    def _feat_opts(self):
        return _lib.FeatOpts(
            kind=_lib.FEATURE_KINDS['filterbank'],
            use_energy=int(self.use_energy), energy_floor=self.energy_floor,
            raw_energy=int(self.raw_energy), htk_compat=int(self.htk_compat),
            use_log_fbank=int(self.use_log_fbank),
            use_power=int(self.use_power))
