"""File formats of FeaturesCollection (counterpart of the reference's
test/test_features_serializers.py): every format round-trips data, times,
dtypes and properties; the Kaldi tables have the documented byte layout."""

import os
import struct

import numpy as np
import pytest

from shennong_b200 import Features, FeaturesCollection, serializers


def make_collection(dtype=np.float32, times_1d=False):
    rng = np.random.default_rng(0)
    coll = FeaturesCollection()
    for i, (n, d) in enumerate([(7, 3), (1, 3), (12, 3)]):
        start = np.arange(n) * 0.01
        times = start if times_1d else np.vstack((start, start + 0.025)).T
        props = {
            'pipeline': [{'name': 'mfcc', 'columns': [0, d - 1]}],
            'mfcc': {'num_ceps': d, 'dither': np.float32(0.0),
                     'window_type': 'povey', 'use_energy': True},
            'cmvn': {'stats': rng.standard_normal((2, d + 1))}}
        coll[f'utt{i}'] = Features(
            rng.standard_normal((n, d)).astype(dtype), times, props)
    return coll


@pytest.mark.parametrize('name,ext', [
    ('numpy', '.npz'), ('pickle', '.pkl'), ('matlab', '.mat'),
    ('kaldi', '.ark'), ('csv', '')])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('times_1d', [False, True])
def test_round_trip(tmp_path, name, ext, dtype, times_1d):
    coll = make_collection(dtype, times_1d)
    path = str(tmp_path / ('feats' + ext))
    coll.save(path)
    with pytest.raises(IOError):
        coll.save(path)                        # never overwrites
    for serializer in (None, name):
        back = FeaturesCollection.load(path, serializer=serializer)
        assert isinstance(back, FeaturesCollection)
        assert set(back.keys()) == set(coll.keys())
        for key in coll:
            assert back[key].dtype == coll[key].dtype
            assert back[key].times.shape == coll[key].times.shape
            if name == 'csv':                  # text format: 1e-18 rounding
                assert back[key].is_close(coll[key])
            else:
                assert back[key] == coll[key], key
    # without the properties
    path = str(tmp_path / ('bare' + ext))
    coll.save(path, with_properties=False)
    back = FeaturesCollection.load(path)
    for key in coll:
        assert back[key].properties == {}
        assert np.allclose(back[key].data, coll[key].data)


def test_guess_and_errors(tmp_path):
    coll = make_collection()
    assert serializers.supported_extensions().keys() == {
        '.npz', '.mat', '.pkl', '.h5f', '.ark', ''}
    assert serializers.supported_serializers().keys() == {
        'numpy', 'matlab', 'pickle', 'h5features', 'kaldi', 'csv'}
    with pytest.raises(ValueError, match='invalid extension'):
        coll.save(str(tmp_path / 'feats.wav'))
    with pytest.raises(ValueError, match='invalid serializer'):
        coll.save(str(tmp_path / 'feats.npz'), serializer='spam')
    with pytest.raises(ValueError, match='must be shennong'):
        serializers.get_serializer(dict, 'a.npz', None)
    with pytest.raises(IOError, match='not found'):
        FeaturesCollection.load(str(tmp_path / 'missing.npz'))
    with pytest.raises(ValueError, match='extension must be'):
        coll.save(str(tmp_path / 'feats.npz'), serializer='kaldi')
    with pytest.raises(ValueError, match='h5features'):
        coll.save(str(tmp_path / 'feats.h5f'))
    # the extension can be overridden by the serializer name
    coll.save(str(tmp_path / 'feats.data'), serializer='pickle')
    assert FeaturesCollection.load(
        str(tmp_path / 'feats.data'), serializer='pickle') == coll
    # invalid features are refused
    bad = make_collection()
    bad['utt0']._data = bad['utt0'].data[:3]
    with pytest.raises(ValueError, match='not valid'):
        bad.save(str(tmp_path / 'bad.npz'))


def test_kaldi_layout_and_scp(tmp_path):
    """<key> SPACE \\0 B 'DM ' \\4 int32 rows \\4 int32 cols, row-major doubles
    (Kaldi's binary table format, what pykaldi's DoubleMatrixWriter emits)"""
    coll = make_collection()
    root = str(tmp_path / 'feats')
    coll.save(root + '.ark', scp=True)
    for name in ('.ark', '.scp', '.times.ark', '.times.scp',
                 '.properties.json'):
        assert os.path.isfile(root + name)
    raw = open(root + '.ark', 'rb').read()
    head = b'utt0 \0BDM \4' + struct.pack('<i', 7) + b'\4' + struct.pack(
        '<i', 3)
    assert raw.startswith(head)
    first = np.frombuffer(raw, dtype='<f8', count=21, offset=len(head))
    assert np.array_equal(first.reshape(7, 3),
                          coll['utt0'].data.astype(np.float64))
    assert len(raw) == sum(len(k) + 1 + 15 + 8 * f.data.size
                           for k, f in coll.items())
    # the scp index points at the binary marker of each entry
    for (key, matrix), line in zip(
            serializers.read_scp(root + '.scp'), open(root + '.scp')):
        assert line.split()[0] == key
        assert np.array_equal(matrix, coll[key].data.astype(np.float64))
    # float matrices ('FM') are read too
    serializers.write_ark(
        root + '.f32.ark', [('a', np.eye(3)), ('b', np.zeros((0, 4)))],
        dtype=np.float32)
    back = dict(serializers.read_ark(root + '.f32.ark'))
    assert back['a'].dtype == np.float32 and np.array_equal(
        back['a'], np.eye(3))
    assert back['b'].shape == (0, 4)
    with pytest.raises(ValueError, match='binary'):
        open(root + '.txt.ark', 'w').write('utt0  [ 1 2 ]\n')
        list(serializers.read_ark(root + '.txt.ark'))
    # missing side files are reported
    os.remove(root + '.times.ark')
    with pytest.raises(IOError, match='times.ark'):
        FeaturesCollection.load(root + '.ark')


def test_json_numpy_content():
    data = {'a': np.arange(6.0).reshape(2, 3), 'b': np.float32(1.5),
            'c': [1, 'x', {'d': np.int64(3)}]}
    back = serializers.json_loads(serializers.json_dumps(data))
    assert np.array_equal(back['a'], data['a']) and back['a'].shape == (2, 3)
    assert back['b'] == 1.5 and back['c'] == [1, 'x', {'d': 3}]


def test_matlab_keeps_dtype_of_small_matrices(tmp_path):
    """a 1 x 1 (or one-frame, one-column) float32 matrix comes back float32"""
    coll = FeaturesCollection(
        one=Features(np.ones((1, 1), np.float32), np.zeros((1, 2))),
        row=Features(np.ones((1, 4), np.float32), np.zeros((1, 2))),
        col=Features(np.ones((3, 1), np.float32), np.zeros((3, 2))),
        dbl=Features(np.ones((2, 2), np.float64), np.zeros((2, 2))))
    coll.save(str(tmp_path / 'f.mat'))
    back = FeaturesCollection.load(str(tmp_path / 'f.mat'))
    for name, feats in coll.items():
        assert back[name].dtype == feats.dtype, name
        assert back[name].shape == feats.shape, name
    assert back == coll


def test_pickle_is_the_reference_format(tmp_path):
    """the .pkl holds the collection OBJECT under the reference's class
    paths (shennong/serializers.py:333-351): a process that has the
    reference package loads it with a plain pickle.load, and files written
    by the reference load here"""
    import pickle
    import subprocess
    import sys
    import textwrap
    coll = FeaturesCollection(
        a=Features(np.random.rand(5, 3).astype(np.float32),
                   np.arange(5) * 0.01, {'pipeline': [], 'x': {'y': 1}}),
        b=Features(np.zeros((2, 3)), np.zeros((2, 2))))
    path = tmp_path / 'feats.pkl'
    coll.save(str(path))
    assert FeaturesCollection.load(str(path)) == coll
    coll.save(str(tmp_path / 'bare.pkl'), with_properties=False)
    bare = FeaturesCollection.load(str(tmp_path / 'bare.pkl'))
    assert bare['a'].properties == {} and np.array_equal(
        bare['a'].data, coll['a'].data)
    # a stand-in for the reference package: same module paths, same attribute
    # names (features.py:62-67), no knowledge of shennong_b200
    fake = tmp_path / 'fake' / 'shennong'
    fake.mkdir(parents=True)
    (fake / '__init__.py').write_text('')
    (fake / 'features.py').write_text(textwrap.dedent("""
        class Features:
            def __init__(self, data, times, properties=None, validate=True):
                self._data, self._times = data, times
                self._properties = {} if properties is None else properties
    """))
    (fake / 'features_collection.py').write_text(
        'class FeaturesCollection(dict):\n    pass\n')
    script = textwrap.dedent(f"""
        import pickle, sys
        sys.path.insert(0, {str(tmp_path / 'fake')!r})
        coll = pickle.load(open({str(path)!r}, 'rb'))
        assert type(coll).__module__ == 'shennong.features_collection'
        assert type(coll['a']).__module__ == 'shennong.features'
        assert coll['a']._data.shape == (5, 3) and coll['a']._properties['x'] == dict(y=1)
        assert 'shennong_b200' not in sys.modules
        pickle.dump(coll, open({str(tmp_path / 'ref.pkl')!r}, 'wb'))
        print('ok')
    """)
    out = subprocess.run([sys.executable, '-c', script], capture_output=True,
                         text=True)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr
    # ... and what the "reference" wrote loads here as this package's classes
    back = FeaturesCollection.load(str(tmp_path / 'ref.pkl'))
    assert type(back) is FeaturesCollection and back == coll
