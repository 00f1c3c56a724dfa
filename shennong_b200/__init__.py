"""shennong_b200: B200-native engine for shennong's frame-based feature path

Drop-in counterpart of ``shennong`` for spectrogram / filterbank / MFCC / PLP /
energy / Kaldi pitch processors and delta / CMVN / VAD post-processors:
same class names, constructor arguments, ``process(Audio) -> Features``
contract and pipeline configuration, executed by hand-written sm_100a CUDA
kernels (``libsnb.so``, C ABI in ``include/snb.h``).
"""

from shennong_b200.audio import Audio
from shennong_b200.features import Features
from shennong_b200.features_collection import FeaturesCollection
from shennong_b200.utterances import Utterance, Utterances

__all__ = ['Audio', 'Features', 'FeaturesCollection', 'Utterance',
           'Utterances']


def version(type=str):
    """Version of the engine as a string or a tuple of integers"""
    numbers = (0, 1, 0)
    return '.'.join(str(n) for n in numbers) if type is str else numbers
