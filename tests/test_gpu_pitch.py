"""GPU parity of Kaldi pitch extraction / post-processing vs the oracle.

Pitch parity against Kaldi itself is unpinned (oracle/README.md); here the
CUDA path is checked against the oracle restatement.  The Viterbi path is a
discrete decision: float summation order can flip near-ties, so a small
fraction of frames may land on a neighbouring lag (0.5 % pitch step)."""

import numpy as np
import pytest

import oracle
from conftest import synth_utterance
from shennong_b200 import Audio, Features
from shennong_b200.processor import (
    KaldiPitchPostProcessor, KaldiPitchProcessor)

pytestmark = pytest.mark.gpu


def check_pitch(out, ref):
    assert out.shape == ref.shape
    close = np.abs(out[:, 1] - ref[:, 1]) <= 1e-4 * ref[:, 1]
    assert close.mean() >= 0.97, f'only {close.mean():.3f} of frames agree'
    # where the lag agrees the NCCF must agree too
    assert np.abs(out[close, 0] - ref[close, 0]).max() < 2e-3
    # disagreeing frames are at most a few lag steps away
    ratio = out[~close, 1] / ref[~close, 1]
    assert np.all(np.abs(np.log(ratio)) < 0.1) if ratio.size else True


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
    assert out.dtype == np.float32
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
    frame) in one ragged batch"""
    sigs = [synth_utterance(i, n) for i, n in
            enumerate([160000, 48000, 16000, 22713])]
    proc = KaldiPitchProcessor()
    outs = proc._extract([Audio(s, 16000) for s in sigs])
    for sig, out in zip(sigs, outs):
        check_pitch(out, oracle.pitch(sig))
    assert outs[0].shape == (998, 2)


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
    assert np.allclose(out.data, ref, rtol=1e-4, atol=1e-4)
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
