"""Name -> processor class registry and processor instantiation for the pipeline

Counterpart of shennong/pipeline_manager.py restricted to the processors of
the frame-based hot path.  Processors the engine does not provide (bottleneck,
CREPE pitch, UBM, VTLN training: out of scope) are listed so
that configurations naming them fail with an explicit message.
"""

import datetime
import importlib
import re

from shennong_b200.audio import Audio
from shennong_b200.logger import get_logger

_UNAVAILABLE = {
    'bottleneck': 'bottleneck features (numpy DNN)',
    'crepe_pitch': 'CREPE pitch (tensorflow)',
    'crepe_pitch_post': 'CREPE pitch (tensorflow)',
    'ubm': 'diagonal UBM training',
    'vtln': 'VTLN warp estimation (use the `warps` argument of '
            'extract_features to apply known warps)'}


class PipelineManager:
    """Instantiates the processors of a pipeline configuration"""
    valid_features = ['spectrogram', 'filterbank', 'mfcc', 'plp']
    """The main features available, excluding post-processing"""

    valid_processors = {
        'energy': ('processor', 'energy', 'EnergyProcessor'),
        'filterbank': ('processor', 'filterbank', 'FilterbankProcessor'),
        'mfcc': ('processor', 'mfcc', 'MfccProcessor'),
        'kaldi_pitch': ('processor', 'pitch_kaldi', 'KaldiPitchProcessor'),
        'kaldi_pitch_post': (
            'processor', 'pitch_kaldi', 'KaldiPitchPostProcessor'),
        'plp': ('processor', 'plp', 'PlpProcessor'),
        'spectrogram': ('processor', 'spectrogram', 'SpectrogramProcessor'),
        'cmvn': ('postprocessor', 'cmvn', 'CmvnPostProcessor'),
        'delta': ('postprocessor', 'delta', 'DeltaPostProcessor'),
        'sliding_window_cmvn': (
            'postprocessor', 'cmvn', 'SlidingWindowCmvnPostProcessor'),
        'vad': ('postprocessor', 'vad', 'VadPostProcessor')}
    """The processors as a dict {name: (package, module, class)}"""

    def __init__(self, config, utterances,
                 log=get_logger('manager', 'warning')):
        self._config = config
        self._utterances = utterances
        self._warps = {}
        self.log = log
        by_speaker = bool(config.get('cmvn', {}).get('by_speaker', False))
        if by_speaker and not utterances.has_speakers():
            raise ValueError(
                'cmvn normalization by speaker requested '
                'but no speaker information provided')
        self._audio_metadata = {
            audio: Audio.scan(audio)
            for audio in set(u.audio_file for u in utterances)}
        self._check_audio_files()
        self.features = [
            k for k in config.keys() if k in self.valid_features][0]
        proc = self.get_features_processor(next(iter(utterances)))
        self.frame_length = proc.frame_length
        self.frame_shift = proc.frame_shift
        self.ndims = proc.ndims

    config = property(lambda self: self._config)
    utterances = property(lambda self: self._utterances)
    audio_metadata = property(lambda self: self._audio_metadata)

    @property
    def warps(self):
        """VTLN warps of the utterances (optional)"""
        return self._warps

    @warps.setter
    def warps(self, value):
        self._warps = value

    def _check_audio_files(self):
        speakers = ''
        if self.utterances.has_speakers():
            speakers = ' from {} speakers'.format(
                len(set(u.speaker for u in self.utterances)))
        self.log.info(
            'get %s utterances%s in %s audio files, total duration: %s',
            len(self.utterances), speakers, len(self.audio_metadata),
            datetime.timedelta(seconds=self.utterances.duration()))
        if not all(m.nchannels == 1 for m in self.audio_metadata.values()):
            raise ValueError('all audio files are not mono')
        rates = set(m.sample_rate for m in self.audio_metadata.values())
        if len(rates) > 1:
            self.log.warning(
                'several sample rates found in audio files: %s, features '
                'extraction pipeline will work but this may not be a good '
                'idea to work on heterogeneous data',
                ', '.join(str(r) + 'Hz' for r in rates))

    @classmethod
    def get_processor_class(cls, name):
        """The (post)processor class called `name`; ValueError if unknown or
        not provided by this engine"""
        if name in _UNAVAILABLE:
            raise ValueError(
                'processor "{}" is not available in shennong_b200: {}'.format(
                    name, _UNAVAILABLE[name]))
        try:
            package, module, klass = cls.valid_processors[name]
        except KeyError:
            raise ValueError('invalid processor "{}"'.format(name)) from None
        mod = importlib.import_module(f'shennong_b200.{package}.{module}')
        return getattr(mod, klass)

    @classmethod
    def get_processor_params(cls, name):
        """Default parameters of processor `name` as a dict"""
        return cls.get_processor_class(name)().get_params()

    @classmethod
    def get_docstring(cls, processor, param, default):
        """One-line documentation of a processor's parameter"""
        doc = getattr(cls.get_processor_class(processor), param).__doc__ or ''
        doc = re.sub(r'\n\n', '. ', doc)
        doc = re.sub(r'\n', ' ', doc)
        doc = re.sub(r'`', '', doc)
        doc = re.sub(':func:', '', doc)
        doc += '. Default is {}.'.format(default)
        doc = re.sub(r'\.+', '.', doc)
        doc = re.sub(r' +', ' ', doc)
        doc = re.sub(r'\. \.', '.', doc)
        return doc.strip()

    def _sample_rate(self, utterance):
        return self.audio_metadata[utterance.audio_file].sample_rate

    def _configure(self, processor):
        processor.log.setLevel(self.log.getEffectiveLevel())
        return processor

    def get_audio(self, utterance):
        return utterance.load_audio()

    def get_features_processor(self, utterance):
        proc = self.get_processor_class(self.features)(
            **self.config[self.features])
        proc.sample_rate = self._sample_rate(utterance)
        return self._configure(proc)

    def get_energy_processor(self, utterance):
        proc = self.get_processor_class('energy')()
        proc.frame_length = self.frame_length
        proc.frame_shift = self.frame_shift
        proc.sample_rate = self._sample_rate(utterance)
        return self._configure(proc)

    def get_vad_processor(self, _=None):
        return self._configure(
            self.get_processor_class('vad')(**self.config['cmvn']['vad']))

    def get_cmvn_processor(self, utterance=None):
        return self._configure(self.get_processor_class('cmvn')(self.ndims))

    def get_pitch_processor(self, utterance):
        params = {k: v for k, v in self.config['pitch'].items()
                  if k not in ('processor', 'postprocessing')}
        params['sample_rate'] = self._sample_rate(utterance)
        params['frame_shift'] = self.frame_shift
        params['frame_length'] = self.frame_length
        if self.config['pitch']['processor'] != 'kaldi':
            self.get_processor_class('crepe_pitch')  # raises
        return self._configure(
            self.get_processor_class('kaldi_pitch')(**params))

    def get_pitch_post_processor(self, _=None):
        return self._configure(self.get_processor_class('kaldi_pitch_post')(
            **self.config['pitch']['postprocessing']))

    def get_delta_processor(self, _=None):
        return self._configure(
            self.get_processor_class('delta')(**self.config['delta']))

    def get_warp(self, utterance):
        return self.warps.get(utterance.name, 1.0)
