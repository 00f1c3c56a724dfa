"""A dict of Features indexed by utterance name

Counterpart of shennong/features_collection.py; the file formats live in
shennong_b200/serializers.py.
"""

import numpy as np

from shennong_b200.features import Features
from shennong_b200.logger import get_logger
from shennong_b200.serializers import get_serializer


class FeaturesCollection(dict):
    """Handles a collection of Features as a dictionary"""

    def is_valid(self):
        return all(f.is_valid() for f in self.values())

    def is_close(self, other, rtol=1e-5, atol=1e-8):
        if self.keys() != other.keys():
            return False
        return all(self[k].is_close(other[k], rtol=rtol, atol=atol)
                   for k in self.keys())

    def partition(self, index):
        """Splits the collection in sub-collections given {item: class}"""
        missing = set(self.keys()) - set(index.keys())
        if missing:
            raise ValueError(
                'following items are not defined in the partition index: {}'
                .format(', '.join(sorted(missing))))
        parts = {}
        for item, label in index.items():
            if item in self:
                parts.setdefault(label, FeaturesCollection())[item] = \
                    self[item]
        return parts

    def trim(self, vad):
        """Keeps the frames where `vad[name]` is true"""
        if vad.keys() != self.keys():
            raise ValueError('Vad keys are different from this keys.')
        out = FeaturesCollection()
        for name, feats in self.items():
            mask = np.asarray(vad[name])
            if mask.dtype != np.dtype('bool'):
                raise ValueError('Vad arrays must be arrays of bool.')
            if mask.shape[0] != feats.nframes:
                raise ValueError(
                    'Vad arrays length must be equal to the number of frames.')
            out[name] = Features(
                feats.data[mask], feats.times[mask],
                properties=feats.properties)
        return out

    def save(self, filename, serializer=None, with_properties=True,
             log=get_logger('features', 'warning'), **kwargs):
        """Saves the collection to `filename`

        The format is guessed from the extension (.npz, .pkl, .mat, .ark,
        .h5f, or a directory name for csv) unless `serializer` names it
        (see serializers.supported_serializers).  `kwargs` are specific to
        the format (e.g. ``scp=True`` for kaldi, ``compress``).
        Raises IOError if the file exists, ValueError on invalid format.
        """
        get_serializer(self.__class__, str(filename), log, serializer).save(
            self, with_properties=with_properties, **kwargs)

    @classmethod
    def load(cls, filename, serializer=None,
             log=get_logger('features', 'warning')):
        """Loads a collection saved with :meth:`save`"""
        return get_serializer(cls, str(filename), log, serializer).load()
