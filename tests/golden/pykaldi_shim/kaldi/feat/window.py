import math

import numpy as np

from ..matrix import Vector


class FrameExtractionOptions:
    def __init__(self):
        self.samp_freq = 16000.0
        self.frame_shift_ms = 10.0
        self.frame_length_ms = 25.0
        self.dither = 1.0
        self.preemph_coeff = 0.97
        self.remove_dc_offset = True
        self.window_type = 'povey'
        self.round_to_power_of_two = True
        self.blackman_coeff = 0.42
        self.snip_edges = True

    def window_shift(self):
        return int(self.samp_freq * 0.001 * self.frame_shift_ms)

    def window_size(self):
        return int(self.samp_freq * 0.001 * self.frame_length_ms)

    def padded_window_size(self):
        n = self.window_size()
        if not self.round_to_power_of_two:
            return n
        p = 1
        while p < n:
            p *= 2
        return p


def num_frames(num_samples, opts, flush=True):
    shift, size = opts.window_shift(), opts.window_size()
    if opts.snip_edges:
        if num_samples < size:
            return 0
        return 1 + (num_samples - size) // shift
    n = (num_samples + shift // 2) // shift
    if flush:
        return n
    raise NotImplementedError


def first_sample_of_frame(frame, opts):
    shift = opts.window_shift()
    if opts.snip_edges:
        return frame * shift
    midpoint = shift * frame + shift // 2
    return midpoint - opts.window_size() // 2


def dither(window, value):
    raise NotImplementedError('golden vectors are generated with dither=0')


def preemphasize(window, coeff):
    """Preemphasize (feature-window.cc)"""
    if coeff == 0.0:
        return
    x = window.numpy()
    c = np.float32(coeff)
    for i in range(x.shape[0] - 1, 0, -1):
        x[i] = np.float32(x[i] - c * x[i - 1])
    x[0] = np.float32(x[0] - c * x[0])


class FeatureWindowFunction:
    def __init__(self, window):
        self.window = window

    @classmethod
    def from_options(cls, opts):
        n = opts.window_size()
        a = 2.0 * math.pi / (n - 1)
        w = np.zeros(n, dtype=np.float32)
        for i in range(n):
            i_fl = float(i)
            t = opts.window_type
            if t == 'hanning':
                v = 0.5 - 0.5 * math.cos(a * i_fl)
            elif t == 'hamming':
                v = 0.54 - 0.46 * math.cos(a * i_fl)
            elif t == 'povey':
                v = math.pow(0.5 - 0.5 * math.cos(a * i_fl), 0.85)
            elif t == 'rectangular':
                v = 1.0
            elif t == 'blackman':
                v = (opts.blackman_coeff - 0.5 * math.cos(a * i_fl)
                     + (0.5 - opts.blackman_coeff) * math.cos(2 * a * i_fl))
            else:
                raise ValueError(t)
            w[i] = v
        return cls(Vector._view(w))


def extract_window(sample_offset, wave, frame, opts, window_function, window, log_energy_pre_window=None):
    """ExtractWindow: delegated to the reference's own Python restatement (plp.py:212-260)"""
    from shennong.processor.plp import _extract_window
    _extract_window(sample_offset, wave, frame, opts, window_function, window, False)
