"""Host-side API (CPU): containers and parameter handling, following the
reference's own tests (test/test_audio.py, test_features.py,
test_utterances.py, test_frames.py, test_window.py, test/processor/*)."""

import numpy as np
import pytest
import scipy.io.wavfile

from shennong_b200 import (
    Audio, Features, FeaturesCollection, Utterance, Utterances)
from shennong_b200.frames import Frames
from shennong_b200.postprocessor import (
    CmvnPostProcessor, DeltaPostProcessor, SlidingWindowCmvnPostProcessor,
    VadPostProcessor)
from shennong_b200.processor import (
    EnergyProcessor, FilterbankProcessor, KaldiPitchPostProcessor,
    KaldiPitchProcessor, MfccProcessor, PlpProcessor, SpectrogramProcessor)
from shennong_b200.window import types, window


# ---- frames (test/test_frames.py:17-187) -------------------------------------
def test_frames_integer_cases():
    f = Frames(sample_rate=1, frame_shift=1, frame_length=1)
    assert f.nframes(10) == 10 and f.samples_per_frame == 1
    assert f.boundaries(3).tolist() == [[0, 1], [1, 2], [2, 3]]
    f = Frames(sample_rate=1, frame_shift=2, frame_length=3)
    assert f.nframes(10) == 4          # 1 + (10 - 3) // 2
    assert f.boundaries(4).tolist() == [[0, 3], [2, 5], [4, 7], [6, 9]]
    assert f.make_frames(np.arange(10)).tolist() == [
        [0, 1, 2], [2, 3, 4], [4, 5, 6], [6, 7, 8]]
    f = Frames(sample_rate=1, frame_shift=2, frame_length=3, snip_edges=False)
    assert f.nframes(10) == 5          # (10 + 1) // 2
    framed = f.make_frames(np.arange(10), writeable=True)
    assert framed.shape == (5, 3) and framed[-1].tolist() == [8, 9, 8]
    f = Frames(sample_rate=16000, frame_shift=0.01, frame_length=0.025)
    assert f.nframes(22713) == 140 and f.nframes(399) == 0
    assert f.times(22713).shape == (140, 2)
    assert f.times(22713)[1].tolist() == [0.01, 0.035]
    with pytest.raises(ValueError):
        Frames(sample_rate=10, frame_shift=0.01).nframes(100)
    view = Frames().make_frames(np.zeros(1000))
    with pytest.raises(ValueError):
        view[0, 0] = 1                 # read-only strided view


# ---- windows (test/test_window.py) -------------------------------------------
def test_windows():
    assert types() == ['blackman', 'hamming', 'hanning', 'povey',
                       'rectangular']
    for t in types():
        w = window(400, t)
        assert w.shape == (400,) and w.max() <= 1.0 and w.min() >= 0
        assert window(1, t).tolist() == [1.0]
    assert window(2, 'povey').tolist() == [1.0, 1.0]
    assert np.allclose(window(2, 'hamming'), [0.08, 0.08])
    with pytest.raises(ValueError):
        window(0)
    with pytest.raises(ValueError):
        window(10, 'foo')


# ---- audio (test/test_audio.py) -----------------------------------------------
def test_audio_astype_and_validation(pcm):
    audio = Audio(pcm, 16000)
    assert audio.nsamples == 22713 and audio.nchannels == 1
    assert audio.duration == pytest.approx(1.4195625)
    assert audio.astype(np.int16) is audio
    f32 = audio.astype(np.float32)
    assert f32.dtype == np.float32 and np.abs(f32.data).max() <= 1
    assert np.array_equal(f32.astype(np.int16).data, pcm)
    i32 = audio.astype(np.int32)
    assert np.array_equal(i32.data, pcm.astype(np.int32) * 2**15)
    assert np.array_equal(i32.astype(np.int16).data, pcm)
    assert np.array_equal(audio.astype(np.float64).astype(np.int16).data, pcm)
    with pytest.raises(ValueError):
        audio.astype(np.int8)
    with pytest.warns(UserWarning):
        with pytest.raises(ValueError):
            Audio(np.ones(10, dtype=np.float32) * 2, 16000)
    assert Audio(pcm.reshape(-1, 1), 16000).data.shape == (22713,)
    stereo = Audio(np.stack([pcm, pcm], 1), 16000)
    assert stereo.nchannels == 2 and stereo.channel(1) == audio
    with pytest.raises(ValueError):
        stereo.channel(2)
    chunks = audio.segment([(0.0, 0.5), (0.5, 1.0)])
    assert [c.nsamples for c in chunks] == [8000, 8000]
    with pytest.raises(ValueError):
        audio.segment([(1.0, 0.5)])
    assert audio.resample(8000).nsamples == 11356


def test_audio_files(tmp_path, pcm):
    path = tmp_path / 'a.wav'
    Audio(pcm, 16000).save(str(path))
    with pytest.raises(ValueError):
        Audio(pcm, 16000).save(str(path))
    meta = Audio.scan(str(path))
    assert (meta.nchannels, meta.sample_rate, meta.nsamples) == (1, 16000, 22713)
    assert Audio.load(str(path)) == Audio(pcm, 16000)
    fpath = tmp_path / 'f.wav'
    scipy.io.wavfile.write(fpath, 16000, (pcm / 2**15).astype(np.float32))
    assert Audio.load(str(fpath)).dtype == np.float32
    assert Audio.scan(str(fpath)).nsamples == 22713
    for bad in ('missing.wav', __file__):
        with pytest.raises(ValueError):
            Audio.scan(str(tmp_path / bad) if bad == 'missing.wav' else bad)
    with pytest.raises(ValueError):
        Audio.load(str(tmp_path / 'missing.wav'))


# ---- features (test/test_features.py) ----------------------------------------
def test_features_container():
    data = np.random.default_rng(0).random((10, 4)).astype(np.float32)
    times = np.vstack((np.arange(10) * 0.01, np.arange(10) * 0.01 + 0.025)).T
    feats = Features(data, times, {'a': 1})
    assert (feats.nframes, feats.ndims, feats.dtype) == (10, 4, np.float32)
    assert feats == feats.copy() and feats.copy() is not feats
    assert feats.copy(dtype=np.float64).dtype == np.float64
    assert feats.copy(subsample=2).shape == (5, 4)
    with pytest.raises(ValueError):
        feats.copy(subsample=0)
    assert feats.is_close(Features(data + 1e-9, times, {'a': 1}))
    assert not feats.is_close(Features(data + 1, times, {'a': 1}))
    assert feats != Features(data, times, {'a': 2})
    for bad in (Features(data, times[:5], validate=False),
                Features(data, times[::-1].copy(), validate=False),
                Features(data * np.nan, times, validate=False),
                Features(data[0], times, validate=False),
                Features(data.tolist(), times, validate=False)):
        assert not bad.is_valid()
    other = Features(data[:, :2] * 2, times.copy(),
                     {'pipeline': [{'name': 'x', 'columns': [0, 1]}], 'x': 3})
    cat = feats.concatenate(other)
    assert cat.shape == (10, 6) and cat.properties['x'] == 3
    assert cat.properties['pipeline'] == [{'name': 'x', 'columns': [4, 5]}]
    short = Features(data[:9, :2], times[:9].copy())
    assert feats.concatenate(short, tolerance=2).shape == (9, 6)
    with pytest.raises(ValueError):
        feats.concatenate(short)
    with pytest.raises(ValueError):
        feats.concatenate(Features(data[:5], times[:5].copy()), tolerance=2)
    with pytest.raises(ValueError):
        feats.concatenate(Features(data, times + 1))
    roundtrip = Features._from_dict(feats._to_dict())
    assert roundtrip == feats
    with pytest.raises(ValueError):
        Features._from_dict({'data': data})


def test_features_collection(tmp_path):
    rng = np.random.default_rng(1)
    coll = FeaturesCollection()
    for i in range(3):
        coll[f'f{i}'] = Features(
            rng.random((5 + i, 3)), np.arange(5 + i) * 0.01, {'i': i})
    assert coll.is_valid() and coll.is_close(coll)
    parts = coll.partition({'f0': 'a', 'f1': 'b', 'f2': 'a'})
    assert sorted(parts) == ['a', 'b'] and sorted(parts['a']) == ['f0', 'f2']
    with pytest.raises(ValueError):
        coll.partition({'f0': 'a'})
    trimmed = coll.trim({k: np.arange(v.nframes) % 2 == 0
                         for k, v in coll.items()})
    assert trimmed['f0'].nframes == 3
    with pytest.raises(ValueError):
        coll.trim({k: np.ones(v.nframes) for k, v in coll.items()})
    for ext in ('.pkl', '.npz'):
        path = str(tmp_path / ('c' + ext))
        coll.save(path)
        assert FeaturesCollection.load(path) == coll
        with pytest.raises(IOError):
            coll.save(path)
    with pytest.raises(ValueError):
        coll.save(str(tmp_path / 'c.h5f'))


# ---- utterances (test/test_utterances.py) ---------------------------------------
def test_utterances(tmp_path, pcm):
    wav = tmp_path / 'w.wav'
    scipy.io.wavfile.write(wav, 16000, pcm)
    w = str(wav)
    utt = Utterance('u1', w, 'spk', 0.1, 0.6)
    assert utt.duration == pytest.approx(0.5) and utt.format == 4
    assert str(utt) == f'u1 {w} spk 0.1 0.6'
    assert utt.load_audio().nsamples == 8000
    with pytest.warns(UserWarning):
        assert Utterance('u', w, 0.5, 10).tstop == pytest.approx(1.4195625)
    for bad in [('u',), ('u', w, 's', 1.0), ('u', w, 's', 0.5, 0.2),
                ('u', w, 'a', 'b'), ('u', 'missing.wav')]:
        with pytest.raises(ValueError):
            Utterance(*bad)
    utts = Utterances([('b', w, 's1'), ('a', w, 's2'), ('c', w, 's1')])
    assert len(utts) == 3 and utts.has_speakers() and utts.format() == 2
    assert utts.format(type=str) == '<utterance-id> <audio-file> <speaker-id>'
    assert [u.name for u in utts] == ['a', 'b', 'c']
    assert sorted(utts.by_speaker()) == ['s1', 's2']
    assert utts.duration() == pytest.approx(3 * 1.4195625)
    path = tmp_path / 'utts.txt'
    utts.save(str(path))
    assert Utterances.load(str(path)) == utts
    for bad in ([], [('a', w), ('b', w, 's')], [('a', w), ('a', w)], [3]):
        with pytest.raises(ValueError):
            Utterances(bad)
    with pytest.raises(ValueError):
        Utterances([('a', w)]).by_speaker()
    fitted = utts.fit_to_duration(1.0)
    assert fitted.duration() == pytest.approx(2.0) and len(fitted) == 2
    with pytest.raises(ValueError):
        utts.fit_to_duration(10.0)
    with pytest.warns(UserWarning):
        utts.fit_to_duration(10.0, truncate=True)


# ---- processors' parameters (test/processor/*.py) ---------------------------------
@pytest.mark.parametrize('cls,nparams', [
    (MfccProcessor, 21), (FilterbankProcessor, 21), (PlpProcessor, 25),
    (SpectrogramProcessor, 12), (EnergyProcessor, 12),
    (KaldiPitchProcessor, 13), (KaldiPitchPostProcessor, 13),
    (DeltaPostProcessor, 2), (VadPostProcessor, 4),
    (SlidingWindowCmvnPostProcessor, 5)])
def test_param_round_trip(cls, nparams):
    proc = cls()
    params = proc.get_params()
    assert len(params) == nparams
    other = cls()
    other.set_params(**params)
    assert other.get_params() == params
    assert cls(**params).get_params() == params
    with pytest.raises(ValueError):
        proc.set_params(foo=1)
    assert repr(proc) == cls.__name__


def test_param_types_and_validation():
    m = MfccProcessor(htk_compat=True, num_bins=20, energy_floor=1.0, dither=2)
    p = m.get_params()
    assert p['htk_compat'] is True and p['num_bins'] == 20
    assert p['energy_floor'] == 1.0 and p['dither'] == 2
    for key in ('sample_rate', 'frame_shift', 'frame_length', 'dither',
                'preemph_coeff', 'blackman_coeff', 'low_freq', 'high_freq',
                'vtln_low', 'vtln_high'):
        assert isinstance(p[key], np.float32), key
    assert m.frame_shift == np.float32(0.01) and m.ndims == 13
    m.set_params(sample_rate=0)
    assert m.get_params()['sample_rate'] == 0
    with pytest.raises(ValueError):
        m.window_type = 'foo'
    plp = PlpProcessor()
    assert isinstance(plp.compress_factor, np.float32)
    for bad in (0, 14):
        with pytest.raises(ValueError):
            plp.num_ceps = bad
    assert FilterbankProcessor(use_energy=True).ndims == 24
    assert SpectrogramProcessor().ndims == 257
    assert SpectrogramProcessor(sample_rate=8000).ndims == 129
    assert SpectrogramProcessor(round_to_power_of_two=False).ndims == 201
    with pytest.raises(ValueError):
        EnergyProcessor(compression='foo')
    pitch = KaldiPitchProcessor()
    assert isinstance(pitch.penalty_factor, np.float32) and pitch.ndims == 2
    assert pitch.frame_shift == 0.01 and pitch.sample_rate == 16000.0
    post = KaldiPitchPostProcessor(add_raw_log_pitch=True)
    assert post.ndims == 4
    vad = VadPostProcessor()
    assert isinstance(vad.energy_threshold, np.float32)
    for kw in ({'energy_mean_scale': -1}, {'frames_context': -1},
               {'proportion_threshold': 0}, {'proportion_threshold': 1}):
        with pytest.raises(ValueError):
            VadPostProcessor(**kw)
    for dim in (0, -1, 1.5, 'a'):
        with pytest.raises(ValueError):
            CmvnPostProcessor(dim)
    with pytest.raises(ValueError):
        CmvnPostProcessor(3, stats=np.zeros((2, 3)))
    cmvn = CmvnPostProcessor(3, stats=np.ones((2, 4)))
    assert cmvn.count == 1.0 and cmvn.ndims == 3
    props = MfccProcessor().get_properties(vtln_warp=1.0)
    assert props['pipeline'] == [{'name': 'mfcc', 'columns': [0, 12]}]
    assert props['mfcc']['vtln_warp'] == 1.0


def test_rasta_filter_frame_by_frame_equals_whole_signal():
    """RastaFilter (public helper of plp.py) against the whole-signal form of
    the same filter (rasta_py), as test/processor/test_plp.py:94-124 does"""
    import scipy.signal
    from shennong_b200.processor.plp import RastaFilter
    taps = -np.arange(-2, 3) / 10.0
    rng = np.random.default_rng(3)
    n = 795
    data = np.stack([np.sin(2 * np.pi * np.arange(n) * 200 / 16000),
                     rng.random(n), np.eye(1, n)[0]], axis=1)
    rasta = RastaFilter(3)
    framewise = np.array([rasta.filter(row, do_log=False) for row in data])
    whole = np.zeros_like(data)
    for c in range(3):
        col = data[:, c]
        zi = scipy.signal.lfilter_zi(taps, 1)
        _, zi = scipy.signal.lfilter(taps, 1, col[:4], zi=zi * col[0])
        whole[4:, c], _ = scipy.signal.lfilter(taps, [1, -0.94], col[4:], zi=zi)
    assert np.array_equal(framewise, whole)
    # log domain round trip: the four priming frames come out as exp(0)
    rasta.reset()
    out = np.array([rasta.filter(row + 1.0) for row in data])
    assert np.all(out[:4] == 1.0) and np.isfinite(out).all()


def test_package_helpers_and_placeholders(capsys, tmp_path, pcm):
    """Small API surface the reference's own tests exercise (test_base.py,
    test_utils.py, test_utterances.py, test_pipeline.py:215-220)"""
    import os
    import scipy.io.wavfile
    import shennong_b200
    from shennong_b200 import Utterance, logger, pipeline, utils
    assert '.'.join(shennong_b200.version(type=tuple)) == shennong_b200.version()
    assert len(shennong_b200.version(type='tuple', full=False)) <= len(
        shennong_b200.version(type=tuple, full=True))
    with pytest.raises(ValueError, match='version type must be str or tuple'):
        shennong_b200.version(type=int)
    assert shennong_b200.version() in shennong_b200.version_long()
    assert 'shennong' in shennong_b200.url()

    @utils.CatchExceptions
    def fails():
        raise ValueError('foo')
    with pytest.raises(SystemExit):
        fails()
    assert 'fatal error: foo' in capsys.readouterr().err

    @utils.CatchExceptions
    def interrupted():
        raise KeyboardInterrupt
    with pytest.raises(SystemExit):
        interrupted()
    assert 'keyboard interruption' in capsys.readouterr().err

    wav = str(tmp_path / 'a.wav')
    scipy.io.wavfile.write(wav, 16000, pcm)
    for args in ((0, wav, None, 1), (0, wav, 0, None)):
        with pytest.raises(ValueError, match='both tstart and tstop'):
            Utterance(*args)
    with pytest.raises(ValueError, match='cannot cast tstart as float'):
        Utterance(0, wav, 'abc', 0)

    os.environ.pop('OMP_NUM_THREADS', None)
    pipeline._check_environment(2, log=logger.get_logger('test', 'info'))
    assert ('working on 2 threads but implicit parallelism is active'
            in capsys.readouterr().err)

    from shennong_b200.processor import BottleneckProcessor, VtlnProcessor
    for cls in (BottleneckProcessor, VtlnProcessor):
        with pytest.raises(NotImplementedError, match='not part of'):
            cls()


def test_compat_alias_runs_shennong_imports():
    """shennong_b200.compat: code written against `shennong` imports resolves
    to the engine; out-of-scope modules are stubs that fail when used"""
    import subprocess
    import sys
    code = '''
import shennong_b200.compat as compat
assert compat.install()
import shennong, shennong_b200
assert shennong is shennong_b200
from shennong.processor.mfcc import MfccProcessor
from shennong.postprocessor.cmvn import CmvnPostProcessor
from shennong import pipeline, Features, Audio
import shennong_b200.processor.mfcc as real
assert MfccProcessor is real.MfccProcessor
assert 'mfcc' in pipeline.valid_features()
from shennong.alignment import AlignmentCollection
try:
    AlignmentCollection.load('x')
except NotImplementedError:
    pass
else:
    raise SystemExit('stub did not raise')
compat.uninstall()
import importlib
try:
    importlib.import_module('shennong.audio')
except ModuleNotFoundError:
    print('ok')
'''
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, '-c', code], capture_output=True,
                         text=True, cwd=root, timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == 'ok', (
        res.stdout + res.stderr)


def test_bench_corpus_generator_is_the_tests_generator():
    """both arms of bench.py (and the parity check inside it) read utterances
    of the same seeded generator the tests use"""
    import importlib.util
    import os
    from conftest import ROOT, synth_utterance
    spec = importlib.util.spec_from_file_location(
        'snb_bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for index in (0, 7, 2047):
        assert np.array_equal(bench.synth_utterance(index, 16000),
                              synth_utterance(index, 16000))
    block = bench.synth_host(3, 2)
    assert np.array_equal(block[:bench.UTT_SAMPLES], bench.synth_utterance(3))
