"""Audio container: waveform samples + sample rate

Counterpart of shennong/audio.py restricted to what the feature hot path and
its callers need: construction, WAV load/scan/save, dtype conversion with the
reference's scaling rules, channel selection, segmentation and (scipy)
resampling.  Decoding of compressed formats (flac, mp3 through pydub/ffmpeg in
the reference) is out of scope: ffmpeg is not part of this engine.
"""

import collections
import functools
import os
import warnings
import wave

import numpy as np
import scipy.io.wavfile
import scipy.signal

_SUPPORTED = (np.int16, np.int32, np.float32, np.float64)


class Audio:
    """An audio signal `data` sampled at `sample_rate` Hz

    `data` is shaped [nsamples] or [nsamples, nchannels] and typed int16,
    int32, float32 or float64 (floats in [-1, 1]).  With `validate` True a
    ValueError is raised for unsupported types or out of range samples
    (shennong/audio.py:112-118, 442-467).
    """
    _metadata = collections.namedtuple(
        '_metadata', 'nchannels sample_rate nsamples duration')

    def __init__(self, data, sample_rate, validate=True):
        self._sample_rate = int(sample_rate)
        if data.ndim > 1 and data.shape[1] == 1:
            data = data[:, 0]
        self._data = data
        if validate and not self.is_valid():
            raise ValueError(f'invalid audio data for type {self.dtype}')

    def __eq__(self, other):
        return (self.sample_rate == other.sample_rate
                and np.array_equal(self.data, other.data))

    @property
    def data(self):
        """The samples as a numpy array"""
        return self._data

    @property
    def sample_rate(self):
        """Sampling frequency in Hertz"""
        return self._sample_rate

    @property
    def nsamples(self):
        return self.data.shape[0]

    @property
    def nchannels(self):
        return 1 if self.data.ndim == 1 else self.data.shape[1]

    @property
    def duration(self):
        """Duration in seconds"""
        return self.nsamples / self.sample_rate

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def precision(self):
        """Bits per sample"""
        return self.dtype.itemsize * 8

    # -- files ---------------------------------------------------------------
    @classmethod
    @functools.lru_cache(maxsize=None)
    def scan(cls, filename):
        """Metadata (nchannels, sample_rate, nsamples, duration) of a WAV file

        Raises ValueError if the file is missing or cannot be parsed.
        """
        filename = str(filename)
        if not os.path.isfile(filename):
            raise ValueError(f'{filename}: file not found')
        try:
            with wave.open(filename, 'r') as wav:
                n, rate = wav.getnframes(), wav.getframerate()
                return cls._metadata(wav.getnchannels(), rate, n, n / rate)
        except (wave.Error, EOFError):
            pass
        try:  # float32 WAVs are not handled by the wave module
            rate, data = scipy.io.wavfile.read(filename, mmap=True)
            nchannels = 1 if data.ndim == 1 else data.shape[1]
            return cls._metadata(
                nchannels, int(rate), data.shape[0], data.shape[0] / rate)
        except Exception:
            raise ValueError(f'cannot scan audio file {filename}') from None

    @classmethod
    def wav_layout(cls, filename):
        """(data offset in bytes, nsamples, sample rate) of a mono 16-bit PCM
        WAV file, None for anything else (compressed, float, multi-channel, malformed):
        the batch entry points read such payloads straight into the pinned
        staging buffer the GPU copies from (shennong_b200.stream.AudioSource)
        instead of going through ``load`` (reference: audio.py:243-286)"""
        try:
            with open(str(filename), 'rb') as fh:
                head = fh.read(12)
                if head[:4] != b'RIFF' or head[8:12] != b'WAVE':
                    return None
                fmt = None
                while True:
                    chunk = fh.read(8)
                    if len(chunk) < 8:
                        return None
                    size = int.from_bytes(chunk[4:8], 'little')
                    if chunk[:4] == b'fmt ':
                        fmt = fh.read(size + (size & 1))[:16]
                    elif chunk[:4] == b'data':
                        if fmt is None or len(fmt) < 16:
                            return None
                        tag = int.from_bytes(fmt[0:2], 'little')
                        nchannels = int.from_bytes(fmt[2:4], 'little')
                        rate = int.from_bytes(fmt[4:8], 'little')
                        bits = int.from_bytes(fmt[14:16], 'little')
                        if tag != 1 or nchannels != 1 or bits != 16:
                            return None
                        offset = fh.tell()
                        left = os.path.getsize(str(filename)) - offset
                        return offset, min(size, left) // 2, rate
                    else:
                        fh.seek(size + (size & 1), 1)
        except OSError:
            return None

    @classmethod
    @functools.lru_cache(maxsize=2)
    def load(cls, filename):
        """Loads a WAV file (int16, int32 or float32 samples)"""
        filename = str(filename)
        if not os.path.isfile(filename):
            raise ValueError(f'{filename}: file not found')
        try:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                rate, data = scipy.io.wavfile.read(filename)
        except Exception as err:
            raise ValueError(
                f'{filename}: cannot read file, {err}') from None
        return cls(data, rate, validate=False)

    def save(self, filename):
        """Writes the signal as a WAV file, refusing to overwrite"""
        filename = str(filename)
        if os.path.isfile(filename):
            raise ValueError(f'{filename}: file already exists')
        if '.' not in os.path.basename(filename):
            raise ValueError(
                f'{filename}: cannot write audio file without extension')
        if filename.split('.')[-1].lower() != 'wav':
            raise ValueError(
                f'{filename}: cannot write file, only the wav format is '
                'supported (flac/mp3 need ffmpeg, out of scope)')
        try:
            scipy.io.wavfile.write(filename, self.sample_rate, self.data)
        except Exception as err:
            raise ValueError(f'{filename}: cannot write file, {err}') from None

    # -- transformations -----------------------------------------------------
    def channel(self, index):
        """Mono signal made of channel `index`"""
        if index >= self.nchannels or index < 0:
            raise ValueError(
                'channel {} does not exist, signal has {} channels'.format(
                    index, self.nchannels))
        if self.nchannels == 1:
            return self
        return Audio(self.data[:, index], self.sample_rate, validate=False)

    def resample(self, sample_rate, backend='sox'):
        """Resampled signal

        Same call surface as the reference (shennong/audio.py:358-423):
        `backend` is 'sox' or 'scipy'.  The sox binary is not available to this
        engine, so both names run the FFT method of scipy.signal -- what the
        reference itself falls back to without sox.  One more backend exists
        here: 'kaldi' converts mono int16 audio on the GPU with Kaldi's
        LinearResample (``snb_resample_batch``; whole batches:
        :func:`shennong_b200.engine.resample_packed`).
        """
        if backend not in ('sox', 'scipy', 'kaldi'):
            raise ValueError(f'backend must be sox or scipy, it is {backend}')
        if sample_rate == self.sample_rate:
            return self
        if backend == 'kaldi':
            return self._resample_device(sample_rate)
        try:
            nsamples = int(self.nsamples * sample_rate / self.sample_rate)
            if sample_rate <= 0 or nsamples <= 0:
                raise ValueError('no sample left')
            with warnings.catch_warnings():
                warnings.simplefilter('ignore', category=FutureWarning)
                data = scipy.signal.resample(self.data, nsamples)
        except (ValueError, ZeroDivisionError):
            raise ValueError(f'resampling at {sample_rate} failed!') from None
        return Audio(data.astype(self.dtype), sample_rate, validate=False)

    def _resample_device(self, sample_rate):
        from shennong_b200 import engine
        if self.nchannels != 1 or self.dtype != np.int16:
            raise ValueError(
                'the kaldi backend resamples mono int16 audio (use '
                'astype(np.int16) and channel())')
        try:
            packed = engine.resample_packed(
                engine.PackedAudio([self.data]), int(self.sample_rate),
                int(sample_rate))
        except (ValueError, RuntimeError) as err:
            if 'CUDA device' in str(err):
                raise
            raise ValueError(f'resampling at {sample_rate} failed!') from None
        n = int(packed.lengths[0])
        return Audio(packed.dev[:n].cpu().numpy(), sample_rate,
                     validate=False)

    @staticmethod
    def _is_valid_dtype(dtype):
        return np.dtype(dtype) in [np.dtype(t) for t in _SUPPORTED]

    def is_valid(self):
        """True when dtype is supported and samples are within its range"""
        if not self._is_valid_dtype(self.dtype):
            warnings.warn(f'unsupported audio data type: {self.dtype}')
            return False
        if self.dtype == np.int16:
            low, high = -2**15, 2**15 - 1
        elif self.dtype == np.int32:
            low, high = -2**31, 2**31 - 1
        else:
            low, high = -1, 1
        if self.data.size and (self.data.min() < low or self.data.max() > high):
            warnings.warn(
                f'invalid audio for type {self.dtype}: boundaries must be in '
                f'({low}, {high}) but are '
                f'({self.data.min()}, {self.data.max()})')
            return False
        return True

    def astype(self, dtype):
        """The signal converted to `dtype` with the reference's scaling

        int16 <-> int32 by 2**15, int16 -> float by 1/2**15, int32 -> float
        by 1/2**30, float -> int16 by 2**15 (C truncation), float -> int32 by
        2**30 (shennong/audio.py:469-518).
        """
        target = np.dtype(dtype)
        if self.dtype == target:
            return self
        if not self._is_valid_dtype(target):
            raise ValueError(f'unsupported audio data type: {dtype}')
        is_float = target.kind == 'f'
        if self.dtype == np.int16:
            data = (self.data / 2**15 if is_float
                    else self.data.astype(np.int32) * 2**15)
        elif self.dtype == np.int32:
            data = self.data / 2**30 if is_float else self.data / 2**15
        else:
            if target == np.int16:
                data = self.data * 2**15
            elif target == np.int32:
                data = self.data * 2**30
            else:
                data = self.data
        return Audio(data.astype(target), self.sample_rate, validate=False)

    def segment(self, segments):
        """List of Audio chunks for a list of (tstart, tstop) pairs in seconds"""
        if not isinstance(segments, list):
            raise ValueError('segments must be a list')
        for seg in segments:
            try:
                if len(seg) != 2:
                    raise ValueError('segments elements must be pairs')
            except TypeError:
                raise ValueError('segments elements must be pairs') from None
            if seg[0] >= seg[1]:
                raise ValueError('time indices in segments must be sorted')
        return [
            Audio(self.data[int(t0 * self.sample_rate):
                            int(t1 * self.sample_rate)],
                  self.sample_rate, validate=False)
            for t0, t1 in segments]
