"""A dict of Features indexed by utterance name

Counterpart of shennong/features_collection.py.  Serialisation supports the
formats that need no third party backend here: pickle (.pkl) and numpy
(.npz); the other formats of the reference (h5features, matlab, kaldi ark,
csv) are "next" rows of the scope table.
"""

import os
import pickle

import numpy as np

from shennong_b200.features import Features


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

    @staticmethod
    def _format(filename, serializer):
        if serializer is None:
            ext = os.path.splitext(filename)[1]
            serializer = {'.pkl': 'pickle', '.npz': 'numpy'}.get(ext)
            if serializer is None:
                raise ValueError(
                    f'invalid extension {ext} of file {filename}, must be '
                    f'.pkl or .npz (other formats are not implemented)')
        if serializer not in ('pickle', 'numpy'):
            raise ValueError(
                f'invalid serializer {serializer}, must be pickle or numpy')
        return serializer

    def save(self, filename, serializer=None, with_properties=True):
        """Saves the collection to `filename` (.pkl or .npz)"""
        filename = str(filename)
        serializer = self._format(filename, serializer)
        if os.path.exists(filename):
            raise IOError(f'file already exists: {filename}')
        if serializer == 'pickle':
            with open(filename, 'wb') as stream:
                pickle.dump(
                    {k: v._to_dict(with_properties) for k, v in self.items()},
                    stream, protocol=4)
        else:
            np.savez_compressed(filename, features=np.asarray(
                {k: v._to_dict(with_properties) for k, v in self.items()},
                dtype=object))

    @classmethod
    def load(cls, filename, serializer=None):
        """Loads a collection saved with :meth:`save`"""
        filename = str(filename)
        serializer = cls._format(filename, serializer)
        if not os.path.isfile(filename):
            raise IOError(f'file not found: {filename}')
        if serializer == 'pickle':
            with open(filename, 'rb') as stream:
                raw = pickle.load(stream)
        else:
            raw = np.load(filename, allow_pickle=True)['features'].item()
        return cls({k: Features._from_dict(v, validate=False)
                    for k, v in raw.items()})
