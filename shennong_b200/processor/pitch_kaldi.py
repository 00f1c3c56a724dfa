"""Kaldi pitch extraction and post-processing
(counterpart of shennong/processor/pitch_kaldi.py)

    Audio --> KaldiPitchProcessor --> KaldiPitchPostProcessor --> Features

Raw pitch is [nframes, 2] (NCCF, pitch in Hz); post-processed pitch is 1 to 4
columns (POV, normalised log-pitch, delta log-pitch, raw log-pitch).
"""

import copy

import numpy as np

from shennong_b200 import _lib, engine
from shennong_b200.base import Option, f32_py, ms_load, ms_store
from shennong_b200.features import Features
from shennong_b200.processor.base import (
    FeaturesPostProcessor, FeaturesProcessor)


class KaldiPitchProcessor(FeaturesProcessor):
    """Extracts the (NCCF, pitch) per frame from a speech signal"""
    sample_rate = Option(
        'Waveform sample frequency in Hertz\n\n'
        'Must match the sample rate of the signal specified in `process`',
        **f32_py())
    frame_shift = Option('Frame shift in seconds', store=ms_store,
                         load=ms_load)
    frame_length = Option('Frame length in seconds', store=ms_store,
                          load=ms_load)
    min_f0 = Option('Minimum F0 to search for in Hertz', **f32_py())
    max_f0 = Option('Maximum F0 to search for in Hertz', **f32_py())
    soft_min_f0 = Option(
        'Minimum F0 to search, applied in soft way, in Hertz\n\n'
        'Must not exceed `min_f0`', **f32_py())
    penalty_factor = Option('Cost factor for F0 change', store=np.float32,
                            load=np.float32)
    lowpass_cutoff = Option('Cutoff frequency for low-pass filter, in Hertz',
                            **f32_py())
    resample_freq = Option(
        'Frequency that we down-sample the signal to, in Hertz\n\n'
        'Must be more than twice `lowpass_cutoff`', **f32_py())
    delta_pitch = Option(
        'Smallest relative change in pitch that the algorithm measures',
        store=np.float32, load=np.float32)
    nccf_ballast = Option(
        'Increasing this factor reduces NCCF for quiet frames\n\n'
        'This helps ensuring pitch continuity in unvoiced regions',
        **f32_py())
    lowpass_filter_width = Option(
        'Integer that determines filter width of lowpass filter\n\n'
        'More gives sharper filter', store=int)
    upsample_filter_width = Option(
        'Integer that determines filter width when upsampling NCCF',
        store=int)

    def __init__(self, sample_rate=16000, frame_shift=0.01,
                 frame_length=0.025, min_f0=50, max_f0=400,
                 soft_min_f0=10, penalty_factor=0.1,
                 lowpass_cutoff=1000, resample_freq=4000,
                 delta_pitch=0.005, nccf_ballast=7000,
                 lowpass_filter_width=1, upsample_filter_width=5):
        super().__init__()
        self.sample_rate = sample_rate
        self.frame_shift = frame_shift
        self.frame_length = frame_length
        self.min_f0 = min_f0
        self.max_f0 = max_f0
        self.soft_min_f0 = soft_min_f0
        self.penalty_factor = penalty_factor
        self.lowpass_cutoff = lowpass_cutoff
        self.resample_freq = resample_freq
        self.delta_pitch = delta_pitch
        self.nccf_ballast = nccf_ballast
        self.lowpass_filter_width = lowpass_filter_width
        self.upsample_filter_width = upsample_filter_width

    @property
    def name(self):
        return 'pitch'

    @property
    def ndims(self):
        return 2

    def times(self, nframes):
        """(tstart, tstop) of the rows returned by `process`"""
        start = np.arange(nframes) * self.frame_shift
        return np.vstack((start, start + self.frame_length)).T

    def _pitch_opts(self):
        o = self.__dict__['_options']
        return _lib.PitchOpts(
            o['sample_rate'], o['frame_shift'], o['frame_length'], 0.0,
            o['min_f0'], o['max_f0'], o['soft_min_f0'], o['penalty_factor'],
            o['lowpass_cutoff'], o['resample_freq'], o['delta_pitch'],
            o['nccf_ballast'], int(o['lowpass_filter_width']),
            int(o['upsample_filter_width']), 1)

    def _check(self, signal):
        if signal.nchannels != 1:
            raise ValueError(
                'audio signal must have one channel, but it has {}'
                .format(signal.nchannels))
        if self.sample_rate != signal.sample_rate:
            raise ValueError(
                'processor and signal mismatch in sample rates: '
                '{} != {}'.format(self.sample_rate, signal.sample_rate))

    def _extract(self, signals):
        for signal in signals:
            self._check(signal)
        pcms = [s.astype(np.int16).data for s in signals]
        plan = engine.pitch_plan(self._pitch_opts())
        packed = engine.PackedAudio(pcms)
        batch = engine.Batch(plan, packed)
        host = engine.to_host(engine.compute_pitch(plan, batch))
        offs = batch.frame_offsets
        return [host[offs[i]:offs[i + 1]] for i in range(len(signals))]

    def _wrap(self, data):
        return Features(
            data, self.times(data.shape[0]),
            properties=self.get_properties())

    def process(self, signal):
        """(NCCF, pitch) of a mono `signal`, float32 [nframes, 2]"""
        return self._wrap(self._extract([signal])[0])

    def _process_batch(self, audios):
        return [self._wrap(d) for d in self._extract(audios)]


class KaldiPitchPostProcessor(FeaturesPostProcessor):
    """Turns raw (NCCF, pitch) into features usable for speech processing"""
    pitch_scale = Option(
        'Scaling factor for the final normalized log-pitch value',
        **f32_py())
    pov_scale = Option(
        'Scaling factor for final probability of voicing feature',
        **f32_py())
    pov_offset = Option(
        'This can be used to add an offset to the POV feature\n\n'
        "Intended for use in Kaldi's online decoding as a substitute for "
        'CMV (cepstral mean normalization)', **f32_py())
    delta_pitch_scale = Option(
        'Term to scale the final delta log-pitch feature', **f32_py())
    delta_pitch_noise_stddev = Option(
        'Standard deviation for noise we add to the delta log-pitch\n\n'
        'The stddev is added before scaling. Should be about the same as '
        'delta-pitch option to pitch creation. The purpose is to get rid of '
        'peaks in the delta-pitch caused by discretization of pitch values.',
        store=np.float32, load=np.float32)
    normalization_left_context = Option(
        'Left-context (in frames) for moving window normalization',
        store=int)
    normalization_right_context = Option(
        'Right-context (in frames) for moving window normalization',
        store=int)
    delta_window = Option(
        'Number of frames on each side of central frame', store=int)
    delay = Option(
        'Number of frames by which the pitch information is delayed',
        store=int)
    add_pov_feature = Option(
        'If true, the warped NCCF is added to output features', store=bool)
    add_normalized_log_pitch = Option(
        'If true, the normalized log-pitch is added to output features\n\n'
        'Normalization is done with POV-weighted mean subtraction over 1.5 '
        'second window.', store=bool)
    add_delta_pitch = Option(
        'If true, time derivative of log-pitch is added to output features',
        store=bool)
    add_raw_log_pitch = Option(
        'If true, time derivative of log-pitch is added to output features',
        store=bool)

    def __init__(self, pitch_scale=2.0, pov_scale=2.0, pov_offset=0.0,
                 delta_pitch_scale=10.0, delta_pitch_noise_stddev=0.005,
                 normalization_left_context=75,
                 normalization_right_context=75,
                 delta_window=2, delay=0,
                 add_pov_feature=True, add_normalized_log_pitch=True,
                 add_delta_pitch=True, add_raw_log_pitch=False):
        super().__init__()
        self.pitch_scale = pitch_scale
        self.pov_scale = pov_scale
        self.pov_offset = pov_offset
        self.delta_pitch_scale = delta_pitch_scale
        self.delta_pitch_noise_stddev = delta_pitch_noise_stddev
        self.normalization_left_context = normalization_left_context
        self.normalization_right_context = normalization_right_context
        self.delta_window = delta_window
        self.delay = delay
        self.add_pov_feature = add_pov_feature
        self.add_normalized_log_pitch = add_normalized_log_pitch
        self.add_delta_pitch = add_delta_pitch
        self.add_raw_log_pitch = add_raw_log_pitch

    @property
    def name(self):
        return 'pitch postprocessing'

    @property
    def ndims(self):
        return (self.add_pov_feature + self.add_normalized_log_pitch
                + self.add_delta_pitch + self.add_raw_log_pitch)

    def get_properties(self, features):
        properties = copy.deepcopy(features.properties)
        properties['pitch'][self.name] = self.get_params()
        properties['pipeline'][0]['columns'] = [0, self.ndims - 1]
        return properties

    def _post_opts(self):
        o = self.__dict__['_options']
        return _lib.PitchPostOpts(
            o['pitch_scale'], o['pov_scale'], o['pov_offset'],
            o['delta_pitch_scale'], o['delta_pitch_noise_stddev'],
            o['normalization_left_context'],
            o['normalization_right_context'], o['delta_window'], o['delay'],
            int(o['add_pov_feature']), int(o['add_normalized_log_pitch']),
            int(o['add_delta_pitch']), int(o['add_raw_log_pitch']))

    def _validate(self, ncols):
        if not (self.add_pov_feature or self.add_normalized_log_pitch
                or self.add_delta_pitch or self.add_raw_log_pitch):
            raise ValueError(
                'at least one of the following options must be True: '
                'add_pov_feature, add_normalized_log_pitch, '
                'add_delta_pitch, add_raw_log_pitch')
        if ncols != 2:
            raise ValueError(
                'data shape must be (_, 2), but it is (_, {})'.format(ncols))

    def process(self, raw_pitch):
        """Post-processes the raw pitch [nframes, 2] -> [nframes, 1..4]

        ValueError if `raw_pitch` has not exactly two columns or if all the
        ``add_*`` options are False.  With ``delay != 0`` Kaldi returns
        ``nframes + delay`` rows, which the reference pairs with the
        ``nframes`` input times (pitch_kaldi.py:535-540): the Features
        constructor rejects that there and here (ValueError).  The delayed
        rows themselves are available from ``engine.process_pitch``.
        """
        self._validate(raw_pitch.shape[1])
        x = engine.from_host(raw_pitch.data, np.float32)
        layout = engine.RowLayout([0, x.shape[0]])
        seed = (engine.next_seed()
                if self.delta_pitch_noise_stddev != 0 else 0)
        out = engine.process_pitch(self._post_opts(), x, layout, seed=seed)
        return Features(
            engine.to_host(out), raw_pitch.times,
            properties=self.get_properties(raw_pitch))
