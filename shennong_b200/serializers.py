"""Saves and loads a FeaturesCollection to/from the file formats of the
reference (shennong/serializers.py:20-55): numpy ``.npz``, pickle ``.pkl``,
matlab ``.mat``, Kaldi ``.ark`` (+ ``.scp``), a directory of ``.csv`` /
``.json`` pairs, and h5features ``.h5f`` (only when the ``h5features`` package
is importable, which it is not in this image).

Same entry points as the reference: :func:`supported_extensions`,
:func:`supported_serializers`, :func:`get_serializer` and one
``FeaturesSerializer`` subclass per format with ``save(features,
with_properties=True, **kwargs)`` / ``load()``.  Differences, all on the
implementation side: the Kaldi tables are read and written natively (the
reference goes through pykaldi's table readers, serializers.py:406-505) with
the same on-disk layout -- binary double matrices ``<key> \\0B DM \\4<rows>
\\4<cols> <data>``, a ``.times.ark`` table and a ``.properties.json`` -- and the
JSON side uses a small encoder compatible with ``json_tricks``' ndarray
dictionaries instead of that package.
"""

import abc
import copy
import json
import os
import importlib
import pickle
import struct

import numpy as np

from shennong_b200.features import Features
from shennong_b200.utils import array2list, list_files_with_extension


def supported_extensions():
    """File extensions mapped to their serializer class"""
    return {
        '.npz': NumpySerializer,
        '.mat': MatlabSerializer,
        '.pkl': PickleSerializer,
        '.h5f': H5featuresSerializer,
        '.ark': KaldiSerializer,
        '': CsvSerializer}


def supported_serializers():
    """Serializer names mapped to their class"""
    return {
        'numpy': NumpySerializer,
        'matlab': MatlabSerializer,
        'pickle': PickleSerializer,
        'h5features': H5featuresSerializer,
        'kaldi': KaldiSerializer,
        'csv': CsvSerializer}


def get_serializer(cls, filename, log, serializer=None):
    """The serializer instance for `filename`, from its extension or from the
    `serializer` name.  Raises ValueError when it cannot be guessed or if
    `cls` is not FeaturesCollection (serializers.py:58-109)."""
    if cls.__name__ != 'FeaturesCollection':
        raise ValueError(
            'The `cls` parameter must be shennong.features.FeaturesCollection')
    if serializer is None:
        ext = os.path.splitext(filename)[1]
        try:
            klass = supported_extensions()[ext]
        except KeyError:
            raise ValueError('invalid extension {}, must be in {}'.format(
                ext, list(supported_extensions().keys())))
    else:
        try:
            klass = supported_serializers()[serializer]
        except KeyError:
            raise ValueError('invalid serializer {}, must be in {}'.format(
                serializer, list(supported_serializers().keys())))
    return klass(cls, filename, log)


# -- JSON with numpy content (json_tricks compatible subset) -----------------
class _NumpyJSONEncoder(json.JSONEncoder):
    def default(self, o):
        if isinstance(o, np.ndarray):
            return {'__ndarray__': o.tolist(), 'dtype': str(o.dtype),
                    'shape': list(o.shape), 'Corder': True}
        if isinstance(o, np.generic):
            return o.item()
        return super().default(o)


def _json_hook(obj):
    if '__ndarray__' in obj:
        return np.asarray(
            obj['__ndarray__'], dtype=obj.get('dtype', None)).reshape(
                obj.get('shape', -1))
    return obj


def json_dumps(data, indent=4):
    return json.dumps(data, indent=indent, cls=_NumpyJSONEncoder)


def json_loads(text):
    return json.loads(text, object_hook=_json_hook)


# -- Kaldi binary matrix tables ----------------------------------------------
_ARK_TYPES = {b'FM': np.dtype('<f4'), b'DM': np.dtype('<f8')}


def write_ark(filename, items, scp=None, dtype=np.float64):
    """Writes (key, 2-D array) pairs as a binary Kaldi matrix table

    `dtype` float64 -> ``DM`` (what the reference writes), float32 -> ``FM``.
    When `scp` is a filename, also writes the ``key file:offset`` index.
    """
    dtype = np.dtype(dtype).newbyteorder('<')
    token = {4: b'FM ', 8: b'DM '}[dtype.itemsize]
    index = []
    with open(filename, 'wb') as stream:
        for key, matrix in items:
            matrix = np.ascontiguousarray(np.atleast_2d(matrix), dtype=dtype)
            if ' ' in key or not key:
                raise ValueError(f'invalid key for a Kaldi table: "{key}"')
            stream.write(key.encode() + b' ')
            index.append((key, stream.tell()))
            stream.write(b'\0B' + token)
            stream.write(b'\4' + struct.pack('<i', matrix.shape[0]))
            stream.write(b'\4' + struct.pack('<i', matrix.shape[1]))
            matrix.tofile(stream)
    if scp:
        with open(scp, 'w') as stream:
            for key, offset in index:
                stream.write(f'{key} {filename}:{offset}\n')


def _read_ark_matrix(stream, where):
    if stream.read(2) != b'\0B':
        raise ValueError(f'{where}: only binary Kaldi tables are supported')
    token = stream.read(3)
    if token[:2] not in _ARK_TYPES or token[2:] != b' ':
        raise ValueError(
            f'{where}: unsupported Kaldi object "{token.decode(errors="replace")}"'
            ' (expected a float or double matrix)')
    dtype = _ARK_TYPES[token[:2]]
    dims = []
    for _ in range(2):
        if stream.read(1) != b'\4':
            raise ValueError(f'{where}: corrupted matrix header')
        dims.append(struct.unpack('<i', stream.read(4))[0])
    count = dims[0] * dims[1]
    data = np.fromfile(stream, dtype=dtype, count=count)
    if data.size != count:
        raise ValueError(f'{where}: truncated matrix')
    return data.reshape(dims)


def read_ark(filename):
    """Yields the (key, matrix) pairs of a binary Kaldi matrix table"""
    with open(filename, 'rb') as stream:
        while True:
            key = bytearray()
            while True:
                char = stream.read(1)
                if not char:
                    if key:
                        raise ValueError(f'{filename}: truncated key')
                    return
                if char == b' ':
                    break
                key += char
            yield key.decode(), _read_ark_matrix(stream, filename)


def read_scp(filename):
    """Yields the (key, matrix) pairs referenced by a Kaldi scp index"""
    with open(filename, 'r') as index:
        for line in index:
            line = line.strip()
            if not line:
                continue
            key, where = line.split(None, 1)
            path, offset = where.rsplit(':', 1)
            with open(path, 'rb') as stream:
                stream.seek(int(offset))
                yield key, _read_ark_matrix(stream, path)


# -- serializers -------------------------------------------------------------
class FeaturesSerializer(metaclass=abc.ABCMeta):
    """Base class of the file serializers (serializers.py:112-221)"""
    def __init__(self, cls, filename, log):
        self._features_collection = cls
        self._filename = str(filename)
        self._log = log

    @property
    def filename(self):
        """Name of the file to read or write"""
        return self._filename

    @abc.abstractmethod
    def _save(self, features, with_properties):  # pragma: nocover
        pass

    @abc.abstractmethod
    def _load(self):  # pragma: nocover
        pass

    def _check_save(self):
        if os.path.isfile(self.filename):
            raise IOError(f'file already exists: {self.filename}')

    def _check_load(self):
        if not os.path.isfile(self.filename):
            raise IOError(f'file not found: {self.filename}')
        if not os.access(self.filename, os.R_OK):
            raise IOError(f'file not readable: {self.filename}')

    def save(self, features, with_properties=True, **kwargs):
        """Saves the `features` collection; IOError if the file exists,
        ValueError if they are not a valid FeaturesCollection"""
        self._check_save()
        if not isinstance(features, self._features_collection):
            raise ValueError('features must be {} but are {}'.format(
                self._features_collection.__name__,
                features.__class__.__name__))
        if not features.is_valid():
            raise ValueError('features are not valid')
        self._save(features, with_properties, **kwargs)

    def load(self, **kwargs):
        """Loads the collection; IOError if the file is missing/unreadable,
        ValueError if its content is not valid"""
        self._check_load()
        features = self._load(**kwargs)
        if not features.is_valid():  # pragma: nocover
            raise ValueError(f'features not valid in "{self.filename}"')
        return features

    def _as_dicts(self, features, with_properties):
        return {k: v._to_dict(with_properties=with_properties)
                for k, v in features.items()}


class NumpySerializer(FeaturesSerializer):
    """numpy '.npz' format"""
    def _save(self, features, with_properties, compress=True):
        self._log.info('writing %s', self.filename)
        save = np.savez_compressed if compress else np.savez
        with open(self.filename, 'wb') as stream:
            save(stream, features=np.asarray(
                self._as_dicts(features, with_properties), dtype=object))

    def _load(self):
        self._log.info('loading %s', self.filename)
        raw = np.load(self.filename, allow_pickle=True)['features'].item()
        return self._features_collection(
            (k, Features._from_dict(v, validate=False))
            for k, v in raw.items())


class _ModuleRef:
    """pickles as ``importlib.import_module(name)``"""
    def __init__(self, name):
        self.name = name

    def __reduce__(self):
        return importlib.import_module, (self.name,)


class _ReferencePickler(pickle.Pickler):
    """Pickles the collection OBJECT like the reference does
    (serializers.py:333-351), with the classes recorded under the reference's
    module paths (``shennong.features.Features``,
    ``shennong.features_collection.FeaturesCollection``): the file loads in
    the reference as its own classes, and here through
    :class:`_ReferenceUnpickler`.  ``with_properties=False`` drops the
    properties like the reference's ``_NoPropertiesPickler``."""
    def __init__(self, stream, with_properties, collection_class):
        super().__init__(stream, protocol=4)
        self._with_properties = with_properties
        self._collection_class = collection_class

    def reducer_override(self, obj):
        if obj is Features:
            return getattr, (_ModuleRef('shennong.features'), 'Features')
        if obj is self._collection_class:
            return getattr, (_ModuleRef('shennong.features_collection'),
                             'FeaturesCollection')
        if isinstance(obj, Features):
            return Features, (
                obj.data, obj.times,
                obj.properties if self._with_properties else None, False)
        return NotImplemented


def _import_reference_module(name):
    if name == 'shennong' or name.startswith('shennong.'):
        name = 'shennong_b200' + name[len('shennong'):]
    return importlib.import_module(name)


class _ReferenceUnpickler(pickle.Unpickler):
    """Resolves the reference's module paths to this package (files written
    by the reference or by :class:`_ReferencePickler`), without installing
    the global import alias of shennong_b200.compat"""
    def find_class(self, module, name):
        if (module, name) == ('importlib', 'import_module'):
            return _import_reference_module
        if module == 'shennong' or module.startswith('shennong.'):
            module = 'shennong_b200' + module[len('shennong'):]
        return super().find_class(module, name)


class PickleSerializer(FeaturesSerializer):
    """python pickle '.pkl' format: the FeaturesCollection object itself, as
    the reference writes it (serializers.py:340-351)"""
    def _save(self, features, with_properties):
        self._log.info('writing %s', self.filename)
        with open(self.filename, 'wb') as stream:
            _ReferencePickler(
                stream, with_properties, type(features)).dump(features)

    def _load(self):
        self._log.info('loading %s', self.filename)
        with open(self.filename, 'rb') as stream:
            raw = _ReferenceUnpickler(stream).load()
        if isinstance(raw, dict) and not isinstance(
                raw, self._features_collection) and all(
                    isinstance(v, dict) for v in raw.values()):
            # files written by round 1 of this package: {name: {data, ...}}
            return self._features_collection(
                (k, Features._from_dict(v, validate=False))
                for k, v in raw.items())
        return raw


class MatlabSerializer(FeaturesSerializer):
    """matlab '.mat' format (scipy.io, like serializers.py:250-330)"""
    def _save(self, features, with_properties, compress=True):
        import scipy.io
        self._log.info('writing %s', self.filename)
        scipy.io.savemat(
            self.filename, self._as_dicts(features, with_properties),
            long_field_names=True, appendmat=False, do_compression=compress)

    def _load(self):
        import scipy.io
        self._log.info('loading %s', self.filename)
        raw = scipy.io.loadmat(
            self.filename, appendmat=False, squeeze_me=True, mat_dtype=False,
            struct_as_record=False)
        features = self._features_collection()
        for key, value in raw.items():
            if key in ('__header__', '__version__', '__globals__'):
                continue
            entry = self._plain(value)
            # squeeze_me also squeezes one-frame / one-dimension features (and
            # turns a 1 x 1 matrix into a Python scalar, losing its dtype):
            # those are read again as stored
            data = entry['data']
            if np.ndim(data) < 2:
                stored = scipy.io.loadmat(
                    self.filename, appendmat=False, squeeze_me=False,
                    mat_dtype=False, struct_as_record=False,
                    variable_names=[key])[key][0, 0].data
                data = np.asarray(data, dtype=stored.dtype)
            data = np.atleast_2d(data)
            times = np.atleast_1d(entry['times'])
            if data.shape[0] != times.shape[0]:
                if data.shape[0] == 1 and times.ndim == 1 and (
                        times.shape[0] == 2 and data.shape[1] != 2):
                    times = times.reshape(1, 2)          # one frame
                else:
                    data = data.reshape(times.shape[0], -1)   # one column
            props = self._restore(entry.get('properties', {}))
            features[key] = Features(data, times, props, validate=False)
        return features

    @classmethod
    def _plain(cls, obj):
        """mat_struct objects to nested dictionaries"""
        if hasattr(obj, '_fieldnames'):
            return {name: cls._plain(getattr(obj, name))
                    for name in obj._fieldnames}
        if isinstance(obj, np.ndarray) and obj.dtype == object:
            return [cls._plain(item) for item in obj]
        return obj

    @staticmethod
    def _restore(properties):
        # matlab collapses a one-element list into its element and every
        # list of numbers into an array: rebuild the 'pipeline' list
        if 'pipeline' in properties:
            pipeline = properties['pipeline']
            if not isinstance(pipeline, list):
                pipeline = [pipeline]
            properties['pipeline'] = [array2list(p) for p in pipeline]
        return properties


class H5featuresSerializer(FeaturesSerializer):
    """h5features '.h5f' format -- needs the h5features package"""
    @staticmethod
    def _backend():
        try:
            import h5features
        except ImportError:
            raise ValueError(
                'the h5features serializer requires the "h5features" python '
                'package, which is not installed')
        return h5features

    def _save(self, features, with_properties, compress=True):
        h5features = self._backend()
        self._log.info('writing %s', self.filename)
        data = h5features.Data(
            list(features.keys()),
            [f.times for f in features.values()],
            [f.data for f in features.values()],
            properties=([f.properties for f in features.values()]
                        if with_properties else None))
        with h5features.Writer(
                self.filename, mode='w', chunk_size='auto',
                compression='lzf' if compress else None) as writer:
            writer.write(data, 'features')

    def _load(self):
        h5features = self._backend()
        self._log.info('loading %s', self.filename)
        data = h5features.Reader(self.filename, groupname='features').read()
        features = self._features_collection()
        for n in range(len(data.items())):
            features[data.items()[n]] = Features(
                data.features()[n], data.labels()[n],
                properties=(data.properties()[n]
                            if data.has_properties() else {}),
                validate=False)
        return features


class KaldiSerializer(FeaturesSerializer):
    """Kaldi ark/scp format: ``<root>.ark`` (features as double matrices),
    ``<root>.times.ark`` and ``<root>.properties.json`` (which also records
    the original dtypes), optionally the ``.scp`` indexes"""
    def __init__(self, cls, filename, log):
        super().__init__(cls, filename, log)
        root, ext = os.path.splitext(self.filename)
        if ext != '.ark':
            raise ValueError(
                'when saving to Kaldi ark format, the file extension must be '
                '".ark", it is "{}"'.format(ext))
        self._fileroot = root

    def _save(self, features, with_properties, scp=False):
        for suffix, getter in (
                ('', lambda f: f.data),
                ('.times', lambda f: np.atleast_2d(f.times))):
            ark = self._fileroot + suffix + '.ark'
            index = self._fileroot + suffix + '.scp' if scp else None
            self._log.info('writing %s', ark)
            write_ark(ark, ((k, getter(v)) for k, v in features.items()),
                      scp=index, dtype=np.float64)
        filename = self._fileroot + '.properties.json'
        self._log.info('writing %s', filename)
        data = {k: (copy.deepcopy(v.properties) if with_properties else {})
                for k, v in features.items()}
        for key in data:
            data[key]['__dtype_data__'] = str(features[key].dtype)
            data[key]['__dtype_times__'] = str(features[key].times.dtype)
        with open(filename, 'w') as stream:
            stream.write(json_dumps(data))

    def _load(self):
        filename = self._fileroot + '.properties.json'
        self._log.info('loading %s', filename)
        if not os.path.isfile(filename):
            raise IOError('file not found: {}'.format(filename))
        with open(filename, 'r') as stream:
            properties = json_loads(stream.read())
        ark = self._fileroot + '.times.ark'
        self._log.info('loading %s', ark)
        if not os.path.isfile(ark):
            raise IOError('file not found: {}'.format(ark))
        times = dict(read_ark(ark))
        self._log.info('loading %s', self.filename)
        data = dict(read_ark(self.filename))
        for key, value in times.items():
            # 1-d times were saved as one row (the number of frames tells
            # them from the [1, 2] times of a one-frame utterance)
            if (value.shape[0] == 1 and key in data
                    and value.shape[1] == data[key].shape[0]):
                times[key] = value.reshape(value.shape[1])
        if properties.keys() != data.keys():
            raise ValueError(
                'invalid features: items differ in data and properties')
        if times.keys() != data.keys():
            raise ValueError(
                'invalid features: items differ in data and times')
        return self._features_collection(
            (k, Features(
                data[k].astype(properties[k]['__dtype_data__']),
                times[k].astype(properties[k]['__dtype_times__']),
                properties={name: p for name, p in properties[k].items()
                            if '__dtype_' not in name},
                validate=False))
            for k in data)


class CsvSerializer(FeaturesSerializer):
    """A directory with one ``<name>.csv`` (times then data columns) and one
    ``<name>.json`` (properties) per features"""
    def _check_load(self):
        if not os.path.isdir(self.filename):
            raise IOError(f'directory not found: {self.filename}')

    def _check_save(self):
        if os.path.exists(self.filename):
            raise IOError(f'already exists: {self.filename}')

    def _save(self, features, with_properties):
        os.makedirs(self.filename)
        self._log.info('writing directory "%s"', self.filename)
        for name, feat in features.items():
            times = (feat.times.reshape((feat.nframes, 1))
                     if feat.times.ndim == 1 else feat.times)
            np.savetxt(
                os.path.join(self.filename, name + '.csv'),
                np.hstack((times, feat.data)),
                header=(f'data_dtype = {feat.dtype}, '
                        f'times_dtype = {feat.times.dtype}, '
                        f'features_ndims = {feat.ndims}'),
                comments='# ')
            if with_properties and feat.properties:
                with open(os.path.join(
                        self.filename, name + '.json'), 'w') as stream:
                    stream.write(json_dumps(feat.properties))

    @staticmethod
    def _parse_header(csv_file):
        with open(csv_file, 'r') as stream:
            header = stream.readline().strip()
        try:
            if header[0] != '#':
                raise IndexError
            fields = header.split(', ')
            return (np.dtype(fields[0].split('= ')[1]),
                    np.dtype(fields[1].split('= ')[1]),
                    int(fields[2].split('= ')[1]))
        except (IndexError, TypeError, ValueError):
            raise ValueError(f'failed to parse header from {csv_file}')

    def _load(self):
        self._log.info('loading directory "%s"', self.filename)
        features = self._features_collection()
        for csv in list_files_with_extension(
                self.filename, '.csv', recursive=False):
            data_dtype, times_dtype, ndims = self._parse_header(csv)
            table = np.atleast_2d(np.loadtxt(csv))
            times = table[:, :table.shape[1] - ndims].astype(times_dtype)
            if times.shape[1] == 1:
                times = times.flatten()
            data = table[:, table.shape[1] - ndims:].astype(data_dtype)
            properties = {}
            sidecar = csv[:-len('.csv')] + '.json'
            if os.path.isfile(sidecar):
                with open(sidecar, 'r') as stream:
                    properties = dict(json_loads(stream.read()))
            name = os.path.basename(csv)[:-len('.csv')]
            features[name] = Features(
                data, times, properties=properties, validate=False)
        return features
