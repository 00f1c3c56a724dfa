"""Perceptual linear predictive features
(counterpart of shennong/processor/plp.py, PlpProcessor)

The reference runs this recipe as a per-frame PYTHON loop over pykaldi
primitives (plp.py:510-626); here it is a tail of the fused CUDA kernel.
With ``rasta=True`` the frame-recursive RASTA filter (plp.py:64-146) runs as
its own kernel between the mel energies and the PLP tail.
"""

import numpy as np

from shennong_b200 import _lib
from shennong_b200.base import Option, f32_py
from shennong_b200.features import Features
from shennong_b200.processor.base import MelFeaturesProcessor


class RastaFilter:
    """RASTA band-pass filter applied frame by frame (host side)

    Public helper of the reference (plp.py:64-146, after rastamat / rasta_py):
    a 5-tap FIR differentiator followed by a single pole at 0.94, run on the
    (log) mel energies along time.  The first four frames only prime the
    filter memory and yield zeros (ones after the inverse log).  The product
    path runs the same recursion on the device (``rasta_kernel``); this class
    keeps the reference's Python entry point for scripts that filter frames
    themselves.

    Parameters
    ----------
    size : int
        Dimension of the frames to filter
    """
    def __init__(self, size):
        import scipy.signal
        self._lfilter, self._lfilter_zi = scipy.signal.lfilter, scipy.signal.lfilter_zi
        taps = np.arange(-2, 3)
        self._num = -taps / np.sum(taps ** 2)
        self._den = np.array([1, -0.94])
        self._size = size
        self.reset()

    def reset(self):
        """Forgets the past frames"""
        self._seen = 0
        self._head = []
        zi = self._lfilter_zi(self._num, 1)
        self._memory = zi if self._size == 1 else np.repeat(
            zi[:, None], self._size, axis=1)

    def filter(self, frame, do_log=True):
        """Filters one frame (shape [size]); `do_log` moves to the log domain
        first and back afterwards, else `frame` is taken as log energies"""
        x = frame
        if do_log:
            x = np.log(x + np.finfo(x.dtype).eps)
        if self._seen < 4:
            self._head.append(x)
            y = np.zeros(x.shape)
            if self._seen == 3:
                head = np.asarray(self._head)
                _, self._memory = self._lfilter(
                    self._num, 1, head, zi=self._memory * head[0], axis=0)
        else:
            y, self._memory = self._lfilter(
                self._num, self._den, [x], zi=self._memory, axis=0)
        self._seen += 1
        y = np.atleast_2d(y)[0, :].astype(x.dtype)
        return np.exp(y) if do_log else y


def _check_num_ceps(proc, value):
    value = int(value)
    if value <= 0:
        raise ValueError('num_ceps must be > 0')
    if value > proc.lpc_order + 1:
        raise ValueError(
            'We must have num_ceps <= lpc_order+1, but {} > {}+1'.format(
                value, proc.lpc_order))
    return value


class PlpProcessor(MelFeaturesProcessor):
    """Perceptive linear predictive features"""
    rasta = Option('Whether to do RASTA filtering', store=bool)
    lpc_order = Option('Order of LPC analysis in PLP computation', store=int)
    num_ceps = Option(
        'Number of cepstra in PLP computation (including C0)\n\n'
        'Must be positive and  smaller or equal to `lpc_order` + 1.',
        store=int, check=_check_num_ceps)
    use_energy = Option(
        'Use energy (instead of C0) for zeroth PLP feature', store=bool)
    energy_floor = Option(
        'Floor on energy (absolute, not relative) in PLP computation',
        **f32_py())
    raw_energy = Option(
        'If true, compute energy before preemphasis and windowing',
        store=bool)
    compress_factor = Option('Compression factor in PLP computation',
                             store=np.float32, load=np.float32)
    cepstral_lifter = Option('Constant that controls scaling of PLPs',
                             **f32_py())
    cepstral_scale = Option('Scaling constant in PLP computation',
                            store=np.float32, load=float)
    htk_compat = Option(
        'If True, get closer to HTK PLP features\n\nPut energy or C0 last.'
        '\n\nWarning: Not sufficient to get HTK compatible features (need '
        'to change other parameters)', store=bool)

    def __init__(self, sample_rate=16000, frame_shift=0.01, frame_length=0.025,
                 rasta=False, dither=1.0, preemph_coeff=0.97,
                 remove_dc_offset=True, window_type='povey',
                 round_to_power_of_two=True, blackman_coeff=0.42,
                 snip_edges=True, num_bins=23, low_freq=20, high_freq=0,
                 vtln_low=100, vtln_high=-500, lpc_order=12, num_ceps=13,
                 use_energy=True, energy_floor=0.0, raw_energy=True,
                 compress_factor=1.0/3.0, cepstral_lifter=22,
                 cepstral_scale=1.0, htk_compat=False):
        super().__init__(
            sample_rate=sample_rate, frame_shift=frame_shift,
            frame_length=frame_length, dither=dither,
            preemph_coeff=preemph_coeff, remove_dc_offset=remove_dc_offset,
            window_type=window_type,
            round_to_power_of_two=round_to_power_of_two,
            blackman_coeff=blackman_coeff, snip_edges=snip_edges,
            num_bins=num_bins, low_freq=low_freq, high_freq=high_freq,
            vtln_low=vtln_low, vtln_high=vtln_high)
        self.rasta = rasta
        self.lpc_order = lpc_order
        self.num_ceps = num_ceps
        self.use_energy = use_energy
        self.energy_floor = energy_floor
        self.raw_energy = raw_energy
        self.compress_factor = compress_factor
        self.cepstral_lifter = cepstral_lifter
        self.cepstral_scale = cepstral_scale
        self.htk_compat = htk_compat

    @property
    def name(self):
        return 'plp'

    @property
    def ndims(self):
        return self.num_ceps

    def _feat_opts(self):
        return _lib.FeatOpts(
            kind=_lib.FEATURE_KINDS['plp'], num_ceps=self.num_ceps,
            use_energy=int(self.use_energy), energy_floor=self.energy_floor,
            raw_energy=int(self.raw_energy),
            cepstral_lifter=self.cepstral_lifter,
            htk_compat=int(self.htk_compat), lpc_order=self.lpc_order,
            compress_factor=self.compress_factor,
            cepstral_scale=self.cepstral_scale, rasta=int(self.rasta))

    def _features(self, data, vtln_warp):
        # the reference skips validation for PLP (plp.py:673-676)
        return Features(
            data, self.times(data.shape[0]),
            properties=self.get_properties(vtln_warp=vtln_warp),
            validate=False)
