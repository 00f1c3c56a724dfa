"""GPU parity of Kaldi pitch extraction / post-processing vs the oracle.

Pitch parity against a Kaldi build is unpinned (oracle/README.md: no
executable Kaldi pitch exists offline).  What is pinned here:

* the CUDA path equals the oracle restatement EXACTLY: the Viterbi state of
  every frame (index work: the second column is 1/lag of that state, compared
  bit for bit) and the NCCF column (same dot-product arithmetic, double
  accumulation in index order, rounded once);
* known answers the oracle cannot fake: synthetic harmonic signals must come
  back at their fundamental within one lag step (the lag grid is 0.5 % apart),
  and the post-processing columns must match a float64 numpy evaluation of the
  published formulas (the reference restates the POV one in
  shennong/processor/pitch_crepe.py:246-253)."""

import numpy as np
import pytest

import oracle
from conftest import numpy_process_pitch, synth_utterance
from shennong_b200 import Audio, Features, engine
from shennong_b200.processor import (
    KaldiPitchPostProcessor, KaldiPitchProcessor)

pytestmark = pytest.mark.gpu


def check_pitch(out, ref):
    """bit-exact state sequence and NCCF"""
    assert out.shape == ref.shape
    assert out.dtype == ref.dtype == np.float32
    same = out[:, 1] == ref[:, 1]
    assert same.all(), (
        f'{(~same).sum()} of {len(same)} frames on a different lag, first at '
        f'{np.flatnonzero(~same)[:5]}')
    assert np.array_equal(out[:, 0], ref[:, 0]), (
        f'NCCF differs by {np.abs(out[:, 0] - ref[:, 0]).max():.3e}')


@pytest.mark.parametrize('kwargs', [
    {}, {'min_f0': 60, 'max_f0': 350}, {'frame_shift': 0.02},
    {'frame_shift': 0.02, 'frame_length': 0.05}, {'penalty_factor': 0.3},
    {'nccf_ballast': 1000, 'soft_min_f0': 20},
    # other table shapes: 1040 states (more Viterbi anchors, fewer warps per
    # CTA), 139 states, generic tap count of the upsampling filter
    {'delta_pitch': 0.002}, {'min_f0': 100, 'max_f0': 200},
    {'upsample_filter_width': 7}, {'resample_freq': 3000,
                                   'lowpass_cutoff': 700}])
def test_pitch_test_wav(pcm, kwargs):
    out = KaldiPitchProcessor(**kwargs).process(Audio(pcm, 16000))
    ref = oracle.pitch(pcm, **kwargs)
    check_pitch(out.data, ref)


def test_pitch_shapes_and_errors(pcm):
    audio = Audio(pcm, 16000)
    assert KaldiPitchProcessor().process(audio).shape == (140, 2)
    assert KaldiPitchProcessor(frame_shift=0.02).process(audio).shape[0] == 70
    with pytest.raises(ValueError):
        KaldiPitchProcessor(sample_rate=8000).process(audio)
    p = KaldiPitchProcessor().process(audio)
    assert p.properties['pitch']['min_f0'] == 50
    assert np.array_equal(
        p.times[:, 0], np.arange(140) * 0.01)


def test_pitch_long_and_batch():
    """10 s synthetic utterances (998 frames: two-phase tail, > recompute
    frame) in one ragged batch (the batch is walked by decreasing length:
    rows must come back in the caller's order), with empty and too-short
    utterances in between"""
    lengths = [48000, 160000, 0, 16000, 300, 22713, 160000, 31999]
    sigs = [synth_utterance(i, n) for i, n in enumerate(lengths)]
    proc = KaldiPitchProcessor()
    outs = proc._extract([Audio(s, 16000) for s in sigs])
    for sig, out in zip(sigs, outs):
        check_pitch(out, oracle.pitch(sig))
    assert outs[1].shape == (998, 2)
    assert outs[2].shape == outs[4].shape == (0, 2)


def test_pitch_many_utterances():
    """more utterances than tracker slots of a small launch: the queue hands
    the rest out; every utterance is checked against the oracle"""
    rng = np.random.default_rng(5)
    lengths = rng.integers(4000, 40000, 300)
    sigs = [synth_utterance(100 + i, int(n)) for i, n in enumerate(lengths)]
    outs = KaldiPitchProcessor()._extract([Audio(s, 16000) for s in sigs])
    for i in range(0, 300, 7):
        check_pitch(outs[i], oracle.pitch(sigs[i]))


@pytest.mark.parametrize('f0', [60.0, 83.0, 110.0, 156.0, 220.0, 297.0, 350.0])
def test_known_answer_harmonic(f0):
    """a clean harmonic complex comes back at its fundamental: every frame
    within two steps of the lag grid (delta_pitch = 0.5 %; the NCCF is
    measured at integer lags of the 4 kHz signal and sinc-interpolated), the
    median within one, with NCCF close to 1"""
    t = np.arange(32000) / 16000.0
    x = sum(4000.0 / h * np.sin(2 * np.pi * h * f0 * t + 0.3 * h)
            for h in range(1, 6))
    sig = np.round(x).astype(np.int16)
    out = KaldiPitchProcessor().process(Audio(sig, 16000)).data
    inner = out[5:-5]
    assert np.all(np.abs(inner[:, 1] / f0 - 1.0) < 0.0101), (
        f0, inner[:, 1].min(), inner[:, 1].max())
    assert abs(np.median(inner[:, 1]) / f0 - 1.0) < 0.0051
    assert np.all(inner[:, 0] > 0.95)
    # and the oracle agrees bit for bit
    check_pitch(out, oracle.pitch(sig))


def test_known_answer_chirp():
    """a slow glide 100 -> 200 Hz is followed (continuity + accuracy)"""
    t = np.arange(48000) / 16000.0
    f = 100.0 + 100.0 * t / t[-1]
    phase = 2 * np.pi * np.cumsum(f) / 16000.0
    x = sum(4000.0 / h * np.sin(h * phase) for h in range(1, 5))
    sig = np.round(x).astype(np.int16)
    out = KaldiPitchProcessor().process(Audio(sig, 16000)).data
    # frame centre times -> instantaneous frequency
    centre = (np.arange(out.shape[0]) * 0.01 + 0.0125)
    expect = 100.0 + 100.0 * centre / t[-1]
    inner = slice(5, -5)
    assert np.all(np.abs(out[inner, 1] / expect[inner] - 1.0) < 0.02)


@pytest.mark.parametrize('kwargs', [
    {}, {'add_raw_log_pitch': True}, {'add_pov_feature': False},
    {'normalization_left_context': 10, 'normalization_right_context': 30},
    {'delta_window': 3, 'pitch_scale': 1.0, 'pov_offset': 0.5}])
def test_process_pitch(pcm, kwargs):
    raw = oracle.pitch(pcm)
    times = np.vstack((np.arange(140) * 0.01, np.arange(140) * 0.01 + .025)).T
    feats = Features(raw, times, {'pitch': {}, 'pipeline': [{}]})
    out = KaldiPitchPostProcessor(
        delta_pitch_noise_stddev=0, **kwargs).process(feats)
    ref = oracle.process_pitch(raw, **kwargs)
    assert out.shape == ref.shape
    assert np.allclose(out.data, ref, rtol=1e-5, atol=2e-6)
    # known answer: float64 numpy evaluation of the published formulas
    full = numpy_process_pitch(
        raw, pitch_scale=kwargs.get('pitch_scale', 2.0),
        pov_offset=kwargs.get('pov_offset', 0.0),
        left=kwargs.get('normalization_left_context', 75),
        right=kwargs.get('normalization_right_context', 75),
        delta_window=kwargs.get('delta_window', 2))
    cols = [c for c, on in enumerate([
        kwargs.get('add_pov_feature', True), True, True,
        kwargs.get('add_raw_log_pitch', False)]) if on]
    assert np.allclose(out.data, full[:, cols], rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        KaldiPitchPostProcessor(
            add_pov_feature=False, add_normalized_log_pitch=False,
            add_delta_pitch=False).process(feats)
    with pytest.raises(ValueError):
        KaldiPitchPostProcessor().process(
            Features(np.zeros((140, 3), np.float32), times))
    # default noise: stochastic third column only
    noisy = KaldiPitchPostProcessor().process(feats)
    clean = KaldiPitchPostProcessor(delta_pitch_noise_stddev=0).process(feats)
    assert np.array_equal(noisy.data[:, :2], clean.data[:, :2])
    delta = (noisy.data[:, 2] - clean.data[:, 2]) / 10.0
    assert 0.002 < delta.std() < 0.01


@pytest.mark.parametrize('delay', [1, 7, 200])
def test_process_pitch_delay(pcm, delay):
    """delay = d: F + d rows per utterance, row t = frame max(0, t - d)
    (Kaldi's ProcessPitch; pitch_kaldi.py:325, 434-440)"""
    raws = [oracle.pitch(pcm), oracle.pitch(synth_utterance(3, 30000)),
            np.zeros((0, 2), np.float32), oracle.pitch(pcm[:9000])]
    offs = np.concatenate(([0], np.cumsum([len(r) for r in raws])))
    post = KaldiPitchPostProcessor(
        delay=delay, delta_pitch_noise_stddev=0, add_raw_log_pitch=True)
    x = engine.from_host(np.concatenate(raws), np.float32)
    out = engine.to_host(engine.process_pitch(
        post._post_opts(), x, engine.RowLayout(offs)))
    assert out.shape == (offs[-1] + delay * len(raws), 4)
    for u, raw in enumerate(raws):
        rows = len(raw) + delay if len(raw) else 0
        got = out[offs[u] + u * delay:offs[u] + u * delay + rows]
        ref = oracle.process_pitch(
            raw, delay=delay, add_raw_log_pitch=True) if len(raw) else got
        assert np.allclose(got, ref, rtol=1e-5, atol=2e-6)
        if len(raw):
            undelayed = oracle.process_pitch(raw, add_raw_log_pitch=True)
            assert np.allclose(got[delay:], undelayed, rtol=1e-5, atol=2e-6)
            assert np.allclose(got[:delay], undelayed[0], rtol=1e-5,
                               atol=2e-6)
    # the host API mirrors the reference: data and times disagree -> ValueError
    times = np.vstack((np.arange(140) * 0.01, np.arange(140) * 0.01 + .025)).T
    with pytest.raises(ValueError):
        post.process(Features(raws[0], times, {'pitch': {}, 'pipeline': [{}]}))
