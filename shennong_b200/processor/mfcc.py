"""Mel frequency cepstral coefficients
(counterpart of shennong/processor/mfcc.py)"""

from shennong_b200 import _lib
from shennong_b200.base import Option, f32_py
from shennong_b200.processor.base import MelFeaturesProcessor


class MfccProcessor(MelFeaturesProcessor):
    """Mel Frequency Cepstral Coeficients"""
    num_ceps = Option(
        'Number of cepstra in MFCC computation (including C0)\n\n'
        'Must be smaller of equal to `num_bins`', store=int)
    use_energy = Option('Use energy (instead of C0) in MFCC computation',
                        store=bool)
    energy_floor = Option(
        'Floor on energy (absolute, not relative) in MFCC computation',
        **f32_py())
    raw_energy = Option(
        'If true, compute energy before preemphasis and windowing',
        store=bool)
    cepstral_lifter = Option('Constant that controls scaling of MFCCs',
                             **f32_py())
    htk_compat = Option(
        'If True, get closer to HTK MFCC features\n\n'
        'Put energy or C0 last and use a factor of sqrt(2) on C0.\n\n'
        'Warning: Not sufficient to get HTK compatible features (need to '
        'change other parameters).', store=bool)

    def __init__(self, sample_rate=16000, frame_shift=0.01,
                 frame_length=0.025, dither=1.0, preemph_coeff=0.97,
                 remove_dc_offset=True, window_type='povey',
                 round_to_power_of_two=True, blackman_coeff=0.42,
                 snip_edges=True, num_bins=23, low_freq=20,
                 high_freq=0, vtln_low=100, vtln_high=-500,
                 num_ceps=13, use_energy=True, energy_floor=0.0,
                 raw_energy=True, cepstral_lifter=22.0,
                 htk_compat=False):
        super().__init__(
            sample_rate=sample_rate, frame_shift=frame_shift,
            frame_length=frame_length, dither=dither,
            preemph_coeff=preemph_coeff, remove_dc_offset=remove_dc_offset,
            window_type=window_type,
            round_to_power_of_two=round_to_power_of_two,
            blackman_coeff=blackman_coeff, snip_edges=snip_edges,
            num_bins=num_bins, low_freq=low_freq, high_freq=high_freq,
            vtln_low=vtln_low, vtln_high=vtln_high)
        self.num_ceps = num_ceps
        self.use_energy = use_energy
        self.energy_floor = energy_floor
        self.raw_energy = raw_energy
        self.cepstral_lifter = cepstral_lifter
        self.htk_compat = htk_compat

    @property
    def name(self):
        return 'mfcc'

    @property
    def ndims(self):
        return self.num_ceps

    def _feat_opts(self):
        return _lib.FeatOpts(
            kind=_lib.FEATURE_KINDS['mfcc'], num_ceps=self.num_ceps,
            use_energy=int(self.use_energy), energy_floor=self.energy_floor,
            raw_energy=int(self.raw_energy),
            cepstral_lifter=self.cepstral_lifter,
            htk_compat=int(self.htk_compat))
