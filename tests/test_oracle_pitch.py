"""Known-answer tests of the oracle's Kaldi pitch restatement (CPU).

No executable Kaldi pitch exists offline (oracle/README.md: parity unpinned),
so the restatement is held to answers it cannot fake: the fundamental of
synthetic harmonic signals, and a float64 numpy evaluation of the published
post-processing formulas (POV sigmoid restated by the reference itself in
shennong/processor/pitch_crepe.py:246-253)."""

import numpy as np
import pytest

import oracle
from conftest import numpy_process_pitch, synth_utterance


@pytest.mark.parametrize('f0', [60.0, 110.0, 220.0, 350.0])
def test_oracle_harmonic_f0(f0):
    t = np.arange(32000) / 16000.0
    x = sum(4000.0 / h * np.sin(2 * np.pi * h * f0 * t + 0.3 * h)
            for h in range(1, 6))
    out = oracle.pitch(np.round(x).astype(np.int16))
    assert out.shape == (198, 2)
    inner = out[5:-5]
    assert np.all(np.abs(inner[:, 1] / f0 - 1.0) < 0.0101)
    assert abs(np.median(inner[:, 1]) / f0 - 1.0) < 0.0051
    assert np.all(inner[:, 0] > 0.95)


def test_oracle_frame_counts_and_ranges(pcm):
    """shapes pinned by the reference (test/processor/test_pitch_kaldi.py:
    39-47): 140 frames on test.wav, 70 at a 20 ms shift"""
    out = oracle.pitch(pcm)
    assert out.shape == (140, 2)
    assert oracle.pitch(pcm, frame_shift=0.02).shape == (70, 2)
    assert np.all((out[:, 1] >= 50) & (out[:, 1] <= 400))
    assert np.all(np.abs(out[:, 0]) <= 1.0 + 1e-6)
    assert oracle.pitch(synth_utterance(0)).shape == (998, 2)


@pytest.mark.parametrize('kwargs', [
    {}, {'normalization_left_context': 10, 'normalization_right_context': 30},
    {'delta_window': 3, 'pitch_scale': 1.0, 'pov_offset': 0.5}])
def test_oracle_process_pitch_formulas(pcm, kwargs):
    raw = oracle.pitch(pcm)
    out = oracle.process_pitch(raw, add_raw_log_pitch=True, **kwargs)
    ref = numpy_process_pitch(
        raw, pitch_scale=kwargs.get('pitch_scale', 2.0),
        pov_offset=kwargs.get('pov_offset', 0.0),
        left=kwargs.get('normalization_left_context', 75),
        right=kwargs.get('normalization_right_context', 75),
        delta_window=kwargs.get('delta_window', 2))
    assert np.allclose(out, ref, rtol=1e-4, atol=1e-5)


def test_oracle_process_pitch_delay(pcm):
    raw = oracle.pitch(pcm)
    base = oracle.process_pitch(raw)
    out = oracle.process_pitch(raw, delay=5)
    assert out.shape == (145, 3)
    assert np.array_equal(out[5:], base)
    assert np.array_equal(out[:5], np.repeat(base[:1], 5, axis=0))
