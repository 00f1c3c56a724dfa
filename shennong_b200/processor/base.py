"""Base classes of the features processors

    Audio --> FeaturesProcessor --> Features

Same contract as shennong/processor/base.py: ``process(signal)`` for one
utterance and ``process_all(utterances, njobs)`` for a collection.  The
difference is below the API: frame-based processors hand the whole batch to
ONE fused CUDA launch (``snb_compute_features``) instead of looping frames on
a CPU thread per utterance.
"""

import abc
import concurrent.futures
import copy

import numpy as np

from shennong_b200 import _lib, engine
from shennong_b200.base import (
    BaseProcessor, Option, f32_f32, ms_load_f32, ms_store)
from shennong_b200.features import Features
from shennong_b200.features_collection import FeaturesCollection
from shennong_b200.utils import get_njobs

_WINDOWS = ['hamming', 'hanning', 'povey', 'rectangular', 'blackman']


class FeaturesProcessor(BaseProcessor, metaclass=abc.ABCMeta):
    """Base class of all the features extraction models"""
    @property
    @abc.abstractmethod
    def name(self):
        """Name of the processor"""

    @property
    @abc.abstractmethod
    def ndims(self):
        """Dimension of the output features frames"""

    def get_properties(self, **kwargs):
        """The processor's properties as a dictionary"""
        params = self.get_params()
        params.update(kwargs)
        return {
            'pipeline': [{'name': self.name, 'columns': [0, self.ndims - 1]}],
            self.name: params}

    @abc.abstractmethod
    def process(self, signal):
        """Returns features computed from an input `signal`"""

    def _process_batch(self, audios, **kwargs):
        """Features of a list of Audio; ``kwargs[name]`` is a list aligned
        with `audios`.  Default: one call to :meth:`process` per utterance;
        frame-based processors override it with a single batched launch."""
        return [self.process(a, **{k: v[i] for k, v in kwargs.items()})
                for i, a in enumerate(audios)]

    def process_all(self, utterances, njobs=None, **kwargs):
        """Features of all the `utterances`, as a FeaturesCollection

        `njobs` sizes the host thread pool that loads the audio files (the
        extraction is a batched GPU launch).  Extra ``kwargs`` are dicts
        {utterance name: value} forwarded to ``process``; ValueError if they
        do not match the utterances (shennong/processor/base.py:89-95).
        """
        njobs = get_njobs(njobs, log=self.log)
        names = list(utterances.by_name().keys())
        for key, value in kwargs.items():
            if not isinstance(value, dict):
                raise ValueError(f'argument "{key}" is not a dict')
            if value.keys() != utterances.by_name().keys():
                raise ValueError(
                    f'utterances and "{key}" have different names')
        utts = [utterances[n] for n in names]
        listed = {k: [v[n] for n in names] for k, v in kwargs.items()}
        feats = self._process_stream(utts, njobs, **listed)
        if feats is None:
            # bounded batches: neither the audio nor the device buffers of
            # the whole corpus are alive at once
            feats, step = [], 256
            for b in range(0, len(utts), step):
                part = utts[b:b + step]
                if njobs > 1 and len(part) > 1:
                    with concurrent.futures.ThreadPoolExecutor(njobs) as pool:
                        audios = list(pool.map(lambda u: u.load_audio(), part))
                else:
                    audios = [u.load_audio() for u in part]
                feats += self._process_batch(
                    audios, **{k: v[b:b + step] for k, v in listed.items()})
        return FeaturesCollection(
            (n, f) for n, f in zip(names, feats) if f is not None)

    def _process_stream(self, utts, njobs, **kwargs):
        """Streamed extraction of a list of Utterance (frame-based processors
        override it); None when the processor has no streamed path"""
        return None


class FeaturesPostProcessor(FeaturesProcessor):
    """Base class of all features post-processors (Features -> Features)

    Defined here (and re-exported by shennong_b200.postprocessor.base, where
    the reference has it: shennong/postprocessor/base.py:15-32) so that the
    pitch post-processor, which lives in the processor package like in the
    reference, does not create an import cycle.
    """
    @abc.abstractmethod
    def process(self, features):
        """Returns features post-processed from input `features`"""

    def get_properties(self, features):
        properties = copy.deepcopy(features.properties)
        properties[self.name] = self.get_params()
        properties.setdefault('pipeline', []).append(
            {'name': self.name, 'columns': [0, self.ndims - 1]})
        return properties


def _check_window(_, value):
    if value not in _WINDOWS:
        raise ValueError(
            'window type must be in {}, it is {}'.format(_WINDOWS, value))


def check_signal(processor, signal):
    """The reference's input checks (processor/base.py:411-419)"""
    if signal.nchannels != 1:
        raise ValueError(
            'signal must have one dimension, but it has {}'
            .format(signal.nchannels))
    if processor.sample_rate != signal.sample_rate:
        raise ValueError(
            'processor and signal mismatch in sample rates: '
            '{} != {}'.format(processor.sample_rate, signal.sample_rate))


class FramesProcessor(FeaturesProcessor, metaclass=abc.ABCMeta):
    """Base class of the frame-based processors (Kaldi framing options)"""
    sample_rate = Option(
        'Waveform sample frequency in Hertz\n\n'
        'Must match the sample rate of the signal specified in `process`',
        **f32_f32())
    frame_shift = Option('Frame shift in seconds', store=ms_store,
                         load=ms_load_f32)
    frame_length = Option('Frame length in seconds', store=ms_store,
                          load=ms_load_f32)
    dither = Option('Amount of dithering\n\n0.0 means no dither',
                    **f32_f32())
    preemph_coeff = Option('Coefficient for use in signal preemphasis',
                           **f32_f32())
    remove_dc_offset = Option(
        'If True, subtract mean from waveform on each frame', store=bool)
    window_type = Option(
        "Type of window\n\nMust be 'hamming', 'hanning', 'povey', "
        "'rectangular' or 'blackman'", check=_check_window)
    round_to_power_of_two = Option(
        'If true, round window size to power of two\n\n'
        'This is done by zero-padding input to FFT', store=bool)
    blackman_coeff = Option(
        'Constant coefficient for generalized Blackman window\n\n'
        "Used only if `window_type` is 'blackman'", **f32_f32())
    snip_edges = Option(
        'If true, output only frames that completely fit in the file\n\n'
        'When True the number of frames depends on the `frame_length`. '
        'If False, the number of frames depends only on the `frame_shift`, '
        'and we reflect the data at the ends.', store=bool)

    def __init__(self, sample_rate=16000, frame_shift=0.01,
                 frame_length=0.025, dither=1.0, preemph_coeff=0.97,
                 remove_dc_offset=True, window_type='povey',
                 round_to_power_of_two=True, blackman_coeff=0.42,
                 snip_edges=True):
        super().__init__()
        self.sample_rate = sample_rate
        self.frame_shift = frame_shift
        self.frame_length = frame_length
        self.dither = dither
        self.preemph_coeff = preemph_coeff
        self.remove_dc_offset = remove_dc_offset
        self.window_type = window_type
        self.round_to_power_of_two = round_to_power_of_two
        self.blackman_coeff = blackman_coeff
        self.snip_edges = snip_edges

    def times(self, nframes):
        """(tstart, tstop) of the rows returned by `process`"""
        start = np.arange(nframes) * self.frame_shift
        return np.vstack((start, start + self.frame_length)).T

    # -- native option structs ------------------------------------------------
    _kind = None          # SNB_FEAT_* name, set by the concrete classes

    def _frame_opts(self, **override):
        o = dict(self.__dict__['_options'])
        o.update(override)
        return _lib.make_frame_opts(
            o['sample_rate'], o['frame_shift'], o['frame_length'],
            o['dither'], o['preemph_coeff'], o['remove_dc_offset'],
            o['window_type'], o['round_to_power_of_two'],
            o['blackman_coeff'], o['snip_edges'])

    def _mel_opts(self):
        return None

    def _feat_opts(self):
        raise NotImplementedError  # pragma: nocover

    def _output_float64(self):
        return False

    def _pcm(self, signal):
        """What the kernels read: the reference casts to int16 before calling
        Kaldi (processor/base.py:428)"""
        return signal.astype(np.int16).data, np.int16

    def _extract(self, signals, vtln_warps=None):
        """One fused launch for `signals`; list of float [nframes, ndims]"""
        for signal in signals:
            check_signal(self, signal)
        pcms, dtypes = zip(*(self._pcm(s) for s in signals))
        dtype = np.float32 if any(d != np.int16 for d in dtypes) else np.int16
        if dtype == np.float32:
            pcms = [np.asarray(p, dtype=np.float32) for p in pcms]
        plan = engine.feature_plan(
            self._frame_opts(), self._mel_opts(), self._feat_opts())
        packed = engine.PackedAudio(pcms, dtype=dtype)
        batch = engine.Batch(plan, packed, vtln_warps)
        seed = engine.next_seed() if self.dither != 0 else 0
        out = engine.compute_features(
            plan, batch, seed=seed, float64=self._output_float64())
        host = engine.to_host(out)
        offs = batch.frame_offsets
        return [host[offs[i]:offs[i + 1]] for i in range(len(signals))]


def stream_features(processor, utts, njobs, warps=None, with_warp=True):
    """process_all of a frame-based processor: the utterances are streamed
    through the fused launch in chunks (shennong_b200.stream), sharded over
    the ranks of torch.distributed when initialised (every rank returns the
    whole collection).  None when a signal is not int16 (float audio goes
    through the per-utterance API, which reproduces the reference's cast)."""
    from shennong_b200 import stream
    from shennong_b200.fused import FusedPipeline
    items, lengths, int16 = stream.audio_items(
        utts, sample_rate=processor.sample_rate)
    if not int16:
        return None
    warp_of = warps
    if warps is not None:
        warps = np.asarray(warps, dtype=np.float32)
    data, parts = stream.extract_corpus(
        FusedPipeline(processor), items, lengths, warps=warps, njobs=njobs)
    feats = [None] * len(utts)
    for index, plan, row0, _, _ in parts:
        for j, i in enumerate(index):
            a = row0 + int(plan.foffs[j])
            block = data[a:a + int(plan.valid[j])]
            kwargs = ({'vtln_warp': warp_of[i] if warp_of is not None
                       else 1.0} if with_warp else {})
            feats[i] = Features._deferred(
                block, _Deferred(processor.times, block.shape[0]),
                _Deferred(processor.get_properties, **kwargs))
    return feats


class _Deferred:
    """a call made at first access (timestamps / properties of the Features
    of a batch: tens of thousands of utterances are wrapped without building
    them up front)"""
    __slots__ = ('fun', 'args', 'kwargs')

    def __init__(self, fun, *args, **kwargs):
        self.fun, self.args, self.kwargs = fun, args, kwargs

    def __call__(self):
        return self.fun(*self.args, **self.kwargs)


class MelFeaturesProcessor(FramesProcessor):
    """Base class of the mel-based processors (filterbank, MFCC, PLP)"""
    num_bins = Option(
        'Number of triangular mel-frequency bins\n\n'
        'The minimal number of bins is 3', store=int)
    low_freq = Option('Low cutoff frequency for mel bins in Hertz',
                      **f32_f32())
    high_freq = Option(
        'High cutoff frequency for mel bins in Hertz\n\n'
        'If `high_freq` < 0, offset from the Nyquist frequency', **f32_f32())
    vtln_low = Option(
        'Low inflection point in piecewise linear VTLN warping function\n\n'
        'In Hertz', **f32_f32())
    vtln_high = Option(
        'High inflection point in piecewise linear VTLN warping function\n\n'
        'In Hertz. If `vtln_high` < 0, offset from `high_freq`', **f32_f32())

    def __init__(self, sample_rate=16000, frame_shift=0.01,
                 frame_length=0.025, dither=1.0, preemph_coeff=0.97,
                 remove_dc_offset=True, window_type='povey',
                 round_to_power_of_two=True, blackman_coeff=0.42,
                 snip_edges=True, num_bins=23, low_freq=20,
                 high_freq=0, vtln_low=100, vtln_high=-500):
        super().__init__(
            sample_rate=sample_rate, frame_shift=frame_shift,
            frame_length=frame_length, dither=dither,
            preemph_coeff=preemph_coeff, remove_dc_offset=remove_dc_offset,
            window_type=window_type,
            round_to_power_of_two=round_to_power_of_two,
            blackman_coeff=blackman_coeff, snip_edges=snip_edges)
        self.num_bins = num_bins
        self.low_freq = low_freq
        self.high_freq = high_freq
        self.vtln_low = vtln_low
        self.vtln_high = vtln_high

    def _mel_opts(self):
        o = self.__dict__['_options']
        return _lib.MelOpts(
            int(o['num_bins']), o['low_freq'], o['high_freq'], o['vtln_low'],
            o['vtln_high'])

    def _features(self, data, vtln_warp):
        return Features(
            data, self.times(data.shape[0]),
            properties=self.get_properties(vtln_warp=vtln_warp))

    def process(self, signal, vtln_warp=1.0):
        """Features of a mono `signal`, optionally VTLN-warped

        Raises ValueError if the signal is not mono or its sample rate differs
        from the processor's; RuntimeError for options Kaldi rejects (e.g.
        num_bins < 3), as the reference does at process time.
        """
        data = self._extract([signal], [vtln_warp])[0]
        return self._features(data, vtln_warp)

    def _process_batch(self, audios, vtln_warp=None):
        warps = vtln_warp if vtln_warp is not None else [1.0] * len(audios)
        datas = self._extract(audios, warps)
        return [self._features(d, w) for d, w in zip(datas, warps)]

    def _process_stream(self, utts, njobs, vtln_warp=None):
        return stream_features(self, utts, njobs, warps=vtln_warp)
