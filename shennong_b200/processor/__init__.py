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
