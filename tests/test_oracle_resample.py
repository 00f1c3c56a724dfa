"""The sample-rate converter of the oracle (Kaldi's LinearResample, flushed:
the step before the path) against golden vectors of an
independent port of the same algorithm (tests/golden/make_resample_golden.py:
torchaudio.functional.resample, Hann-windowed sinc, 6 zeros, rolloff 0.99)."""

import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden',
                      'resample_sinc_hann.npz')


@pytest.fixture(scope='module')
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize('pair', [
    (16000, 8000), (16000, 44100), (16000, 11025), (44100, 16000),
    (8000, 16000), (16000, 4000)])
def test_resample_golden(golden, pair):
    want = golden['%d_%d' % pair]
    got = oracle.resample(golden['pcm'], *pair)
    assert got.shape == want.shape and got.dtype == np.float32
    # float32 evaluation against a float64 one: 1e-6 of the amplitude
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()


def test_resample_counts_and_identity():
    x = np.arange(1000, dtype=np.float32)
    # Kaldi: n_in * rate_out / rate_in outputs, the last one dropped when it
    # falls exactly on the end of the signal
    assert len(oracle.resample(x, 16000, 8000)) == 500
    assert len(oracle.resample(x, 16000, 44100)) == 2757
    assert len(oracle.resample(x[:1], 16000, 8000)) == 1
    assert len(oracle.resample(x[:0], 16000, 8000)) == 0
    # a band-limited tone keeps its amplitude and frequency
    t = np.arange(16000) / 16000.0
    tone = 1000 * np.sin(2 * np.pi * 440 * t)
    y = oracle.resample(tone, 16000, 8000)
    want = 1000 * np.sin(2 * np.pi * 440 * np.arange(8000) / 8000.0)
    assert np.abs(y - want)[100:-100].max() < 1.0
