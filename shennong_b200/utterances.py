"""Utterance and Utterances: the index of what to extract

Counterpart of shennong/utterances.py: an utterance is
``<utterance-id> <audio-file> [<speaker-id>] [<tstart> <tstop>]``.
"""

import collections
import os
import random
import warnings

from shennong_b200.audio import Audio

VALID_FORMATS = {
    1: '<utterance-id> <audio-file>',
    2: '<utterance-id> <audio-file> <speaker-id>',
    3: '<utterance-id> <audio-file> <tstart> <tstop>',
    4: '<utterance-id> <audio-file> <speaker-id> <tstart> <tstop>'}


def _as_float(value, what):
    try:
        return float(value)
    except (TypeError, ValueError):
        raise ValueError(f'cannot cast {what} as float: {value}') from None


class Utterance:
    """A single utterance built from 2 to 5 fields (see VALID_FORMATS)"""
    def __init__(self, *args):
        if not 2 <= len(args) <= 5:
            raise ValueError(f'invalid utterance format: {args}')
        self._format = len(args) - 1
        self._name, self._audio = args[0], args[1]
        self._speaker = args[2] if len(args) in (3, 5) else None
        self._tstart = self._tstop = None
        if len(args) >= 4:
            if (args[-2] is None) != (args[-1] is None):
                raise ValueError(
                    'both tstart and tstop must be defined or None, but '
                    f'(tstart, tstop)=({args[-2]}, {args[-1]})')
        if len(args) >= 4 and args[-1] is not None:
            self._tstart = _as_float(args[-2], 'tstart')
            self._tstop = _as_float(args[-1], 'tstop')
            if self._tstart < 0 or self._tstart >= self._tstop:
                raise ValueError(
                    'we must have 0 <= tstart < tstop, but '
                    f'(tstart, tstop)=({self._tstart}, {self._tstop})')
        # scanning raises if the file is missing or invalid
        self._duration = Audio.scan(self._audio).duration
        if self._tstart is not None:
            if self._tstop > self._duration:
                warnings.warn(
                    f'{self._audio}: file duration is {self._duration} but '
                    f'asking interval ({self._tstart}, {self._tstop}), '
                    f'will be truncated')
                self._tstop = self._duration
            self._duration = self._tstop - self._tstart

    name = property(lambda self: self._name)
    audio_file = property(lambda self: self._audio)
    speaker = property(lambda self: self._speaker)
    tstart = property(lambda self: self._tstart)
    tstop = property(lambda self: self._tstop)
    duration = property(lambda self: self._duration)
    format = property(lambda self: self._format)

    def __eq__(self, other):
        return str(self) == str(other)

    def __str__(self):
        fields = [self.name, self.audio_file]
        if self.speaker is not None:
            fields.append(self.speaker)
        if self.tstart is not None:
            fields += [self.tstart, self.tstop]
        return ' '.join(str(f) for f in fields)

    def load_audio(self):
        """The Audio of the utterance (segmented on tstart/tstop if any)"""
        audio = Audio.load(self._audio)
        if self.tstart or self.tstop:
            audio = audio.segment([(self.tstart, self.tstop)])[0]
        return audio


class Utterances:
    """A collection of Utterance with a homogeneous format and unique names"""
    def __init__(self, utterances):
        parsed = []
        for utt in utterances:
            if not isinstance(utt, Utterance):
                try:
                    utt = Utterance(*utt)
                except TypeError:
                    raise ValueError(
                        f'utterance must be an iterable, not {utt}') from None
            parsed.append(utt)
        if not parsed:
            raise ValueError('empty input utterances')
        formats = set(u.format for u in parsed)
        if len(formats) != 1:
            raise ValueError('utterances format is not homogeneous')
        self._format = formats.pop()
        dups = [n for n, c in collections.Counter(
            u.name for u in parsed).items() if c > 1]
        if dups:
            raise ValueError(
                f'duplicates found in utterances: {", ".join(dups)}')
        parsed.sort(key=lambda u: (u.audio_file, u.name))
        self._utterances = {u.name: u for u in parsed}

    def __len__(self):
        return len(self._utterances)

    def __iter__(self):
        return iter(self._utterances.values())

    def __getitem__(self, name):
        return self._utterances[name]

    def __eq__(self, other):
        return self._utterances == other._utterances

    @classmethod
    def load(cls, filename):
        """Loads utterances from a text file, one per line"""
        if not os.path.isfile(filename):
            raise ValueError(f'{filename} not found')
        with open(filename, 'r') as stream:
            lines = [line.strip() for line in stream]
        return cls([line.split(' ') for line in lines if line])

    def save(self, filename):
        with open(filename, 'w') as stream:
            stream.write('\n'.join(str(u) for u in self) + '\n')

    def format(self, type=int):
        """Format code (int) or its description (str)"""
        return VALID_FORMATS[self._format] if type is str else self._format

    def has_speakers(self):
        return self._format in (2, 4)

    def by_speaker(self):
        """{speaker: [Utterance]}; ValueError without speaker information"""
        if not self.has_speakers():
            raise ValueError('utterances have no speaker information')
        out = collections.defaultdict(list)
        for utt in self:
            out[utt.speaker].append(utt)
        return out

    def by_name(self):
        return self._utterances

    def duration(self):
        return sum(u.duration for u in self)

    def fit_to_duration(self, duration, truncate=False, shuffle=False):
        """Subset keeping `duration` seconds per speaker
        (shennong/utterances.py:331-412)"""
        if duration <= 0:
            raise ValueError(
                f'duration must be a positive number, it is {duration}')
        segments = []
        for speaker, utts in self.by_speaker().items():
            if shuffle:
                random.shuffle(utts)
            remaining = duration
            for utt in utts:
                tstart = 0 if utt.tstart is None else utt.tstart
                tstop = (utt.duration - tstart if utt.tstop is None
                         else utt.tstop)
                if utt.duration >= remaining:
                    segments.append(Utterance(
                        utt.name, utt.audio_file, utt.speaker, tstart,
                        tstart + remaining))
                    remaining = 0
                    break
                segments.append(Utterance(
                    utt.name, utt.audio_file, utt.speaker, tstart, tstop))
                remaining -= utt.duration
            if remaining > 0:
                message = (
                    f'speaker {speaker}: only {duration - remaining}s of '
                    f'audio available but {duration}s requested')
                if truncate:
                    warnings.warn(message)
                else:
                    raise ValueError(message)
        return Utterances(segments)
