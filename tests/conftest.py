"""Shared fixtures.  Tests marked `gpu` need a B200; everything else runs on
CPU (oracle vs golden vectors, host logic, C-ABI symbols)."""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'kaldi_compliance.npz')


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: test needs a CUDA device (run with -m gpu on a B200)')


@pytest.fixture(scope='session', autouse=True)
def _build_native():
    """The checker (oracle) and the product library are built on demand"""
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope='session')
def golden():
    data = np.load(GOLDEN)
    manifest = json.loads(bytes(data['manifest']).decode())
    return data, manifest


@pytest.fixture(scope='session')
def pcm(golden):
    """The reference's test/data/test.wav samples (16 kHz int16 mono)"""
    return golden[0]['pcm']


@pytest.fixture(scope='session')
def audio(pcm):
    from shennong_b200 import Audio
    return Audio(pcm, 16000)


def synth_utterance(index, nsamples=160000, sample_rate=16000):
    """Synthetic utterance of BASELINE.md section 3: 5-harmonic tone + noise,
    seeded per utterance"""
    rng = np.random.default_rng(20260925 + index)
    f0 = rng.uniform(80, 300)
    phases = rng.uniform(0, 2 * np.pi, 5)
    t = np.arange(nsamples) / sample_rate
    x = sum(3000.0 / h * np.sin(2 * np.pi * h * f0 * t + phases[h - 1])
            for h in range(1, 6))
    x = x + 500.0 * rng.standard_normal(nsamples)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def scale_close(actual, desired, tol=1e-4):
    """The parity gate of BASELINE.md: max|a-b| <= tol * max|ref| per matrix"""
    actual, desired = np.asarray(actual), np.asarray(desired)
    assert actual.shape == desired.shape, (actual.shape, desired.shape)
    if desired.size == 0:
        return
    scale = max(float(np.abs(desired).max()), 1e-30)
    err = float(np.abs(actual.astype(np.float64) - desired).max())
    assert err <= tol * scale, f'max abs err {err:.3e} > {tol} * {scale:.3e}'
