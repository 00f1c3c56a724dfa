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


__version__ = '0.1.0'


def url():
    """Where the documentation of the reference API lives (the engine keeps it)"""
    return 'https://docs.cognitive-ml.fr/shennong'


def version(type=str, full=False):
    """Version of the engine as a string or a tuple of strings

    Same call surface as ``shennong.version`` (shennong/__init__.py:41-64):
    `type` is ``str``/``tuple`` (or their names), `full` keeps any
    pre-release field after (major, minor, patch).
    """
    if type not in (str, tuple, 'str', 'tuple'):
        raise ValueError(
            'version type must be str or tuple, it is {}'.format(type))
    fields = tuple(__version__.split('.'))
    if not full:
        fields = fields[:3]
    return fields if type in (tuple, 'tuple') else '.'.join(fields)


def version_long():
    """Version, origin and licence note in a few lines"""
    import datetime
    return (
        'shennong_b200-{} (B200-native engine for the shennong feature path)\n'
        'API modelled on shennong (copyright 2018-{} Inria, licence GPL3)\n'
        'see the documentation of that API at {}\n'.format(
            version(), datetime.date.today().year, url()))
