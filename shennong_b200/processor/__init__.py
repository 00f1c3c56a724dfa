"""Features extraction processors (same names as shennong.processor)"""

from shennong_b200.processor.energy import EnergyProcessor
from shennong_b200.processor.filterbank import FilterbankProcessor
from shennong_b200.processor.mfcc import MfccProcessor
from shennong_b200.processor.pitch_kaldi import (
    KaldiPitchProcessor, KaldiPitchPostProcessor)
from shennong_b200.processor.plp import PlpProcessor
from shennong_b200.processor.spectrogram import SpectrogramProcessor

__all__ = [
    'EnergyProcessor', 'FilterbankProcessor', 'MfccProcessor',
    'KaldiPitchProcessor', 'KaldiPitchPostProcessor', 'PlpProcessor',
    'SpectrogramProcessor']


def _out_of_scope(name, what):
    """Importable placeholder for a processor of the reference that is outside
    the frame-based feature path: scripts that import the
    name keep working, building one fails with an explicit message"""
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            f'{name} is not part of shennong_b200 ({what}): the engine covers '
            'the frame-based feature path only')
    return type(name, (), {'__init__': __init__, '__doc__': _out_of_scope.__doc__})


BottleneckProcessor = _out_of_scope(
    'BottleneckProcessor', 'bottleneck DNN features')
CrepePitchProcessor = _out_of_scope(
    'CrepePitchProcessor', 'CREPE pitch, tensorflow')
CrepePitchPostProcessor = _out_of_scope(
    'CrepePitchPostProcessor', 'CREPE pitch, tensorflow')
OneHotProcessor = _out_of_scope('OneHotProcessor', 'alignments / one-hot')
FramedOneHotProcessor = _out_of_scope(
    'FramedOneHotProcessor', 'alignments / one-hot')
VtlnProcessor = _out_of_scope('VtlnProcessor', 'VTLN training, UBM/GMM')
