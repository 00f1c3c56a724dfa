"""Framing arithmetic (counterpart of shennong/frames.py)

Frame counts come from ``snb_num_frames`` -- the same routine the CUDA engine
uses to lay out its batches -- so host bookkeeping and device indexing cannot
disagree.
"""

import numpy as np

from shennong_b200 import _lib
from shennong_b200.base import BaseProcessor, Option, ms_load, ms_store


class Frames(BaseProcessor):
    """Cuts arrays in overlapping frames"""
    sample_rate = Option('Waveform sample frequency in Hertz',
                         store=np.float32, load=float)
    frame_shift = Option('Frame shift in seconds', store=ms_store,
                         load=ms_load)
    frame_length = Option('Frame length in seconds', store=ms_store,
                          load=ms_load)
    snip_edges = Option(
        'If true, output only frames that completely fit in the file',
        store=bool)

    def __init__(self, sample_rate=16000, frame_shift=0.01,
                 frame_length=0.025, snip_edges=True):
        self.sample_rate = sample_rate
        self.frame_shift = frame_shift
        self.frame_length = frame_length
        self.snip_edges = snip_edges

    @property
    def name(self):
        return 'frames'

    def _opts(self):
        o = self.__dict__['_options']
        return _lib.make_frame_opts(
            o['sample_rate'], o['frame_shift'], o['frame_length'], 0.0, 0.0,
            False, 'povey', True, 0.42, o['snip_edges'])

    @property
    def samples_per_frame(self):
        """Number of samples in one frame"""
        return int(self.frame_length * self.sample_rate)

    @property
    def samples_per_shift(self):
        """Number of samples between two frames"""
        return int(self.frame_shift * self.sample_rate)

    def nframes(self, nsamples):
        """Number of frames extracted from `nsamples` samples"""
        if self.samples_per_shift == 0:
            raise ValueError('cannot compute nframes: sample rate too low')
        return int(_lib.lib().snb_num_frames(int(nsamples),
                                             _lib.ref(self._opts())))

    def first_sample_of_frame(self, frame):
        return int(frame * self.samples_per_shift)

    def last_sample_of_frame(self, frame):
        return int(self.first_sample_of_frame(frame) + self.samples_per_frame)

    def times(self, nsamples):
        """(tstart, tstop) of each frame, shape [nframes, 2]"""
        start = np.arange(self.nframes(nsamples)) * self.frame_shift
        return np.vstack((start, start + self.frame_length)).T

    def boundaries(self, nframes):
        """(istart, istop) sample indices of each frame, shape [nframes, 2]"""
        first = np.asarray(
            [self.first_sample_of_frame(i) for i in range(nframes)],
            dtype=int).reshape(nframes, 1)
        return np.hstack((first, first + self.samples_per_frame))

    def make_frames(self, array, writeable=False):
        """`array` cut in frames along its first axis

        With `snip_edges` False the end of the array is mirrored
        (shennong/frames.py:211-215).  Returns a read-only strided view unless
        `writeable` is True.
        """
        nframes = self.nframes(array.shape[0])
        if not self.snip_edges:
            extra = self.last_sample_of_frame(nframes - 1) - array.shape[0]
            array = np.concatenate((array, array[-extra - 1:-1][::-1]))
        length, shift = self.samples_per_frame, self.samples_per_shift
        if writeable:
            out = np.empty((nframes, length) + array.shape[1:], array.dtype)
            for i in range(nframes):
                out[i] = array[i * shift:i * shift + length]
            return out
        return np.lib.stride_tricks.as_strided(
            array, shape=(nframes, length) + array.shape[1:],
            strides=(array.strides[0] * shift,) + array.strides,
            writeable=False)
