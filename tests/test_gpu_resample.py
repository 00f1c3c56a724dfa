"""Sample-rate conversion on the device (snb_resample_batch)
against the CPU oracle (bit-exact: same weights, same accumulation order),
the committed golden vectors, and the host resamplers of the reference's
Audio.resample as a sanity bound."""

import os

import numpy as np
import pytest
import scipy.signal

import oracle
from conftest import synth_utterance
from shennong_b200 import Audio, engine
from shennong_b200.processor import MfccProcessor

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden',
                      'resample_sinc_hann.npz')


@pytest.mark.parametrize('pair', [
    (16000, 8000), (16000, 44100), (16000, 11025), (44100, 16000),
    (8000, 16000), (16000, 4000), (16000, 12345)])
def test_resample_batch_equals_oracle(pair):
    lengths = [16000, 401, 22713, 7, 1, 48000, 9999]
    sigs = [synth_utterance(40 + i, n) for i, n in enumerate(lengths)]
    packed, outf = engine.resample_packed(
        engine.PackedAudio(sigs), *pair, float32=True)
    outf = outf.cpu().numpy()
    out16 = packed.dev.cpu().numpy()
    for sig, start, n in zip(sigs, packed.starts, packed.lengths):
        ref = oracle.resample(sig, *pair)
        assert n == len(ref)
        got = outf[start:start + n]
        assert np.array_equal(got, ref)
        assert np.array_equal(
            out16[start:start + n],
            np.clip(np.trunc(ref), -32768, 32767).astype(np.int16))
    assert np.all(packed.starts % 8 == 0)


def test_resample_golden_vectors():
    g = np.load(GOLDEN)
    for key in g.files:
        if key == 'pcm':
            continue
        pair = tuple(int(v) for v in key.split('_'))
        _, outf = engine.resample_packed(
            engine.PackedAudio([g['pcm']]), *pair, float32=True)
        got = outf.cpu().numpy()[:len(g[key])]
        assert np.abs(got - g[key]).max() <= 1e-6 * np.abs(g[key]).max()


def test_audio_resample_kaldi_backend(pcm):
    audio = Audio(pcm, 16000)
    low = audio.resample(8000, backend='kaldi')
    assert low.sample_rate == 8000 and low.dtype == np.int16
    assert low.nsamples == len(oracle.resample(pcm, 16000, 8000))
    # same signal as the host polyphase / FFT resamplers up to their filters
    poly = scipy.signal.resample_poly(pcm.astype(np.float64), 1, 2)
    n = min(len(poly), low.nsamples)
    err = np.abs(low.data[:n] - poly[:n])[50:-50]
    assert err.max() < 0.02 * np.abs(poly).max()
    host = audio.resample(8000, backend='scipy')
    n = min(host.nsamples, low.nsamples)
    corr = np.corrcoef(host.data[:n].astype(float), low.data[:n].astype(float))
    assert corr[0, 1] > 0.98
    assert audio.resample(16000, backend='kaldi') is audio
    with pytest.raises(ValueError):
        Audio(pcm.astype(np.float32) / 2**15, 16000).resample(
            8000, backend='kaldi')
    with pytest.raises(ValueError):
        audio.resample(8000, backend='ffmpeg')


def test_resampled_batch_feeds_the_feature_kernels():
    """44.1 kHz recordings converted and extracted without leaving the device"""
    sigs = [synth_utterance(60 + i, n) for i, n in enumerate((44100, 30000))]
    packed = engine.resample_packed(engine.PackedAudio(sigs), 44100, 16000)
    proc = MfccProcessor(dither=0)
    plan = engine.feature_plan(
        proc._frame_opts(), proc._mel_opts(), proc._feat_opts())
    batch = engine.Batch(plan, packed)
    feats = engine.compute_features(plan, batch).cpu().numpy()
    offs = batch.frame_offsets
    for i, sig in enumerate(sigs):
        low = np.clip(np.trunc(oracle.resample(sig, 44100, 16000)),
                      -32768, 32767).astype(np.int16)
        want = proc.process(Audio(low, 16000)).data
        assert np.array_equal(feats[offs[i]:offs[i + 1]], want)
