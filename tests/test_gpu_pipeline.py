"""pipeline.extract_features on the GPU: equals the composition of the
individual processors (the reference's schedule, pipeline.py:570-648) and
produces the same properties layout (test/test_pipeline.py:276-420)."""

import numpy as np
import pytest
import scipy.io.wavfile
import scipy.signal

import oracle
from conftest import scale_close, synth_utterance
from shennong_b200 import Audio, Utterances, pipeline
from shennong_b200.postprocessor import (
    CmvnPostProcessor, DeltaPostProcessor, VadPostProcessor)
from shennong_b200.processor import (
    EnergyProcessor, KaldiPitchPostProcessor, KaldiPitchProcessor,
    MfccProcessor)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def corpus(tmp_path_factory, pcm):
    root = tmp_path_factory.mktemp('wavs')
    entries = []
    lengths = [22713, 48000, 16000, 31000]
    for i, n in enumerate(lengths):
        data = pcm if i == 0 else synth_utterance(i, n)
        path = root / f'w{i}.wav'
        scipy.io.wavfile.write(path, 16000, data)
        entries.append((f'utt{i}', str(path), f'spk{i % 2}'))
    return entries


def no_dither(config):
    for key in ('mfcc', 'filterbank', 'plp', 'spectrogram'):
        if key in config:
            config[key]['dither'] = 0
    if 'pitch' in config:
        config['pitch']['postprocessing']['delta_pitch_noise_stddev'] = 0
    return config


@pytest.mark.parametrize('features', ['mfcc', 'filterbank', 'plp',
                                      'spectrogram'])
@pytest.mark.parametrize('with_delta', [False, True])
def test_features_and_delta(corpus, features, with_delta):
    config = no_dither(pipeline.get_default_config(
        features, with_delta=with_delta))
    utts = Utterances([e[:2] for e in corpus[:3]])
    feats = pipeline.extract_features(config, utts, njobs=2)
    assert list(feats.keys()) == ['utt0', 'utt1', 'utt2']
    proc = pipeline.PipelineManager.get_processor_class(features)(
        **config[features])
    for name, path in [e[:2] for e in corpus[:3]]:
        ref = proc.process(Audio.load(path))
        if with_delta:
            ref = DeltaPostProcessor().process(ref)
        got = feats[name]
        assert got.shape == ref.shape
        assert np.allclose(got.data, ref.data, rtol=1e-5, atol=1e-5)
        assert np.array_equal(got.times, ref.times)
        assert got.properties['pipeline'] == ref.properties['pipeline']
        assert set(got.properties) == set(ref.properties) | {'audio'}
        assert got.properties['audio']['sample_rate'] == 16000
    assert feats['utt0'].shape[0] == 140


@pytest.mark.parametrize('by_speaker', [True, False])
@pytest.mark.parametrize('with_vad', [True, False])
def test_cmvn_delta_pitch_full_pipeline(corpus, by_speaker, with_vad):
    config = no_dither(pipeline.get_default_config(
        'mfcc', with_pitch='kaldi', with_cmvn=True, with_delta=True))
    config['cmvn']['by_speaker'] = by_speaker
    config['cmvn']['with_vad'] = with_vad
    utts = Utterances(corpus)
    feats = pipeline.extract_features(config, utts, njobs=2)
    assert set(feats.keys()) == {e[0] for e in corpus}

    # the reference's schedule with the individual processors
    mfcc, energy, vad = (MfccProcessor(dither=0), EnergyProcessor(dither=0),
                         VadPostProcessor())
    base, weights = {}, {}
    for utt in utts:
        audio = utt.load_audio()
        base[utt.name] = mfcc.process(audio)
        if with_vad:
            weights[utt.name] = vad.process(
                energy.process(audio)).data.reshape(-1).astype(np.float32)
    groups = {}
    for utt in utts:
        groups.setdefault(utt.speaker if by_speaker else utt.name,
                          []).append(utt.name)
    for members in groups.values():
        cmvn = CmvnPostProcessor(13)
        for name in members:
            cmvn.accumulate(base[name], weights.get(name))
        for name in members:
            ref = DeltaPostProcessor().process(cmvn.process(base[name]))
            got = feats[name]
            assert got.shape == (ref.shape[0], 42)
            scale_close(got.data[:, :39], ref.data, tol=2e-4)
            assert np.allclose(
                got.properties['cmvn']['stats'], cmvn.stats, rtol=1e-6)
    # pitch columns and properties
    for utt in utts:
        audio = utt.load_audio()
        raw = KaldiPitchProcessor().process(audio)
        post = KaldiPitchPostProcessor(delta_pitch_noise_stddev=0).process(raw)
        got = feats[utt.name]
        assert np.allclose(got.data[:, 39:], post.data, atol=1e-5)
        assert set(got.properties.keys()) == {
            'audio', 'mfcc', 'cmvn', 'pitch', 'delta', 'speaker', 'pipeline'}
        assert got.properties['pipeline'] == [
            {'name': 'mfcc', 'columns': [0, 12]},
            {'name': 'cmvn', 'columns': [0, 12]},
            {'name': 'delta', 'columns': [0, 38]},
            {'name': 'pitch', 'columns': [39, 41]}]
        assert 'pitch postprocessing' in got.properties['pitch']
    # CMVN'd base columns have zero mean per group when VAD is off
    if not with_vad and not by_speaker:
        for f in feats.values():
            assert np.abs(f.data[:, :13].mean(0)).max() < 1e-4


def test_vtln_warps_and_save_load(corpus, tmp_path):
    config = no_dither(pipeline.get_default_config('mfcc'))
    utts = Utterances([e[:3] for e in corpus[:4]])
    by_spk = pipeline.extract_features(
        config, utts, warps={'spk0': 0.9, 'spk1': 1.1})
    by_utt = pipeline.extract_features(
        config, utts, warps={'utt0': 0.9, 'utt1': 1.1, 'utt2': 0.9,
                             'utt3': 1.1})
    assert by_spk == by_utt
    for name, path, spk in [e[:3] for e in corpus[:4]]:
        warp = 0.9 if spk == 'spk0' else 1.1
        ref = oracle.features(
            'mfcc', Audio.load(path).data, vtln_warp=warp)
        scale_close(by_spk[name].data, ref, tol=1e-4)
        assert by_spk[name].properties['mfcc']['vtln_warp'] == warp
    with pytest.raises(ValueError):
        pipeline.extract_features(config, utts, warps={'nobody': 1.0})
    for ext in ('.pkl', '.npz'):
        path = tmp_path / f'feats{ext}'
        by_spk.save(str(path))
        assert type(by_spk).load(str(path)) == by_spk


def test_segments(corpus):
    """<utt> <wav> <spk> <tstart> <tstop> utterances (test_pipeline.py:347+)"""
    path = corpus[0][1]
    utts = Utterances([('a', path, 's', 0.2, 1.2), ('b', path, 's', 0.0, 0.5)])
    config = no_dither(pipeline.get_default_config('mfcc', with_delta=True))
    feats = pipeline.extract_features(config, utts)
    assert feats['a'].shape == (98, 39) and feats['b'].shape == (48, 39)
    assert feats['a'].properties['audio']['tstart'] == 0.2
    assert feats['a'].properties['audio']['tstop'] == 1.2
    assert feats['a'].properties['audio']['duration'] == 1.0
    chunk = Audio.load(path).segment([(0.2, 1.2)])[0]
    ref = oracle.deltas(oracle.features('mfcc', chunk.data))
    scale_close(feats['a'].data, ref, tol=1e-4)


def test_errors(corpus):
    config = pipeline.get_default_config('mfcc', with_cmvn=True)
    with pytest.raises(ValueError, match='no speaker information'):
        pipeline.extract_features(config, Utterances([e[:2] for e in corpus[:2]]))


def test_extract_features_warp_and_sweep(corpus):
    """extract_features_warp (pipeline.py:669-696 of the reference: main
    features with one warp, then deltas, no CMVN) and its batched sweep over
    a grid of warps: equal to the per-utterance processors, bit-identical
    between the single-warp and the sweep forms"""
    config = no_dither(pipeline.get_default_config('mfcc', with_delta=True))
    utts = Utterances(corpus[:3])
    grid = [0.85, 1.0, 1.2]
    sweep = pipeline.extract_features_warp_sweep(config, utts, grid)
    assert sorted(sweep) == grid
    proc = MfccProcessor(**config['mfcc'])
    for warp in grid:
        single = pipeline.extract_features_warp(config, utts, warp)
        assert list(single.keys()) == ['utt0', 'utt1', 'utt2']
        for name, path, _ in corpus[:3]:
            ref = DeltaPostProcessor().process(
                proc.process(Audio.load(path), vtln_warp=warp))
            got = single[name]
            assert got.shape == ref.shape
            assert np.array_equal(got.times, ref.times)
            assert np.allclose(got.data, ref.data, rtol=1e-5, atol=1e-5)
            assert got.properties['mfcc']['vtln_warp'] == warp
            assert np.array_equal(sweep[warp][name].data, got.data)
    # warping changes the features, the identity warp does not
    base = pipeline.extract_features(config, utts)
    assert np.array_equal(sweep[1.0]['utt0'].data, base['utt0'].data)
    assert not np.allclose(sweep[0.85]['utt0'].data, base['utt0'].data)
    with pytest.raises(ValueError):
        pipeline.extract_features_warp(
            no_dither(pipeline.get_default_config(
                'spectrogram', with_cmvn=False)), utts, 1.1)


def test_extract_features_full_mixed_rates(pcm, tmp_path):
    """The reference's "difficult case" (test/test_pipeline.py:347-420):
    different sampling rates, float32 audio, segments, and a speaker whose
    utterances have different rates -- its CMVN statistics are pooled"""
    import warnings
    from shennong_b200 import FeaturesCollection
    wav = str(tmp_path / 'test.wav')
    wav_f32 = str(tmp_path / 'test.float32.wav')
    wav_8k = str(tmp_path / 'test.8k.wav')
    scipy.io.wavfile.write(wav, 16000, pcm)
    scipy.io.wavfile.write(wav_f32, 16000, (pcm / 32768.0).astype(np.float32))
    scipy.io.wavfile.write(
        wav_8k, 8000, np.round(scipy.signal.resample_poly(
            pcm.astype(np.float64), 1, 2)).astype(np.int16))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')        # u3 is longer than its file
        index = Utterances([
            ('u1', wav, 's1', 0, 1),
            ('u2', wav_f32, 's2', 1, 1.2),
            ('u3', wav_8k, 's1', 1, 3)])
    config = pipeline.get_default_config(
        'mfcc', with_cmvn=True, with_delta=True, with_pitch='kaldi')
    config['cmvn']['with_vad'] = False
    feats = pipeline.extract_features(config, index, njobs=2)
    for utt in ('u1', 'u2', 'u3'):
        assert feats[utt].dtype == np.float32
    p1, p2, p3 = (feats[u].properties for u in ('u1', 'u2', 'u3'))
    assert p1['audio']['duration'] == 1.0
    assert p2['audio']['duration'] == pytest.approx(0.2)
    assert p3['audio']['duration'] < 0.5
    assert p1['mfcc'] == p2['mfcc']
    assert p1['mfcc']['sample_rate'] != p3['mfcc']['sample_rate']
    assert p1.keys() == {
        'audio', 'mfcc', 'cmvn', 'pitch', 'delta', 'speaker', 'pipeline'}
    assert p1.keys() == p2.keys() == p3.keys()
    assert p1['pipeline'] == p2['pipeline'] == p3['pipeline']
    assert feats['u1'].shape == (98, 42)
    assert feats['u2'].shape == (18, 42)
    assert feats['u3'].shape == (40, 42)
    assert feats['u2'].data[:, :13].mean() == pytest.approx(0.0, abs=1e-5)
    assert feats['u2'].data[:, :13].std() == pytest.approx(1.0, abs=1e-5)
    data = np.vstack((feats['u1'].data[:, :13], feats['u3'].data[:, :13]))
    assert data.mean() == pytest.approx(0.0, abs=1e-5)
    assert data.std() == pytest.approx(1.0, abs=1e-5)
    assert np.abs(data.mean()) <= np.abs(feats['u1'].data[:, :13].mean())
    assert np.abs(data.std() - 1.0) <= np.abs(
        feats['u1'].data[:, :13].std() - 1.0)
    assert np.abs(data.mean()) <= np.abs(feats['u3'].data[:, :13].mean())
    assert np.abs(data.std() - 1.0) <= np.abs(
        feats['u3'].data[:, :13].std() - 1.0)
    filename = str(tmp_path / 'feats.npz')
    feats.save(filename)
    assert FeaturesCollection.load(filename) == feats
