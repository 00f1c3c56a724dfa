"""GPU parity of the post-processors (delta, CMVN, sliding CMN, VAD)"""

import numpy as np
import pytest

import oracle
from conftest import scale_close
from shennong_b200 import Features, FeaturesCollection
from shennong_b200.postprocessor import (
    CmvnPostProcessor, DeltaPostProcessor, SlidingWindowCmvnPostProcessor,
    VadPostProcessor, apply_cmvn)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mfcc(pcm):
    data = oracle.features('mfcc', pcm)
    times = np.vstack((np.arange(140) * 0.01, np.arange(140) * 0.01 + 0.025)).T
    return Features(data, times, {'pipeline': []})


@pytest.mark.parametrize('order,window', [(0, 1), (1, 2), (2, 2), (3, 1),
                                          (2, 5), (1, 30)])
def test_delta(mfcc, order, window):
    out = DeltaPostProcessor(order=order, window=window).process(mfcc)
    ref = oracle.deltas(mfcc.data, order, window)
    assert out.shape == (140, 13 * (order + 1))
    assert np.allclose(out.data, ref, rtol=1e-5, atol=1e-5)
    assert np.array_equal(out.data[:, :13], mfcc.data)    # test_delta.py:27-35
    assert out.properties['delta'] == {'order': order, 'window': window}


def test_delta_short_and_invalid(mfcc):
    for n in (1, 2, 3):
        feats = Features(mfcc.data[:n].copy(), mfcc.times[:n].copy())
        out = DeltaPostProcessor().process(feats)
        assert np.allclose(out.data, oracle.deltas(feats.data), atol=1e-5)
    with pytest.raises(ValueError):
        DeltaPostProcessor(window=0)
    with pytest.raises(ValueError):
        DeltaPostProcessor(window=1000)
    with pytest.raises(ValueError):
        DeltaPostProcessor().ndims


@pytest.mark.parametrize('norm_vars', [True, False])
def test_cmvn(mfcc, norm_vars):
    before = mfcc.data.copy()
    proc = CmvnPostProcessor(13)
    assert proc.count == 0
    with pytest.raises(ValueError):
        proc.process(mfcc)
    proc.accumulate(mfcc)
    ref_stats = oracle.cmvn_accumulate(mfcc.data)
    assert proc.count == 140
    assert np.allclose(proc.stats, ref_stats, rtol=1e-12)
    out = proc.process(mfcc, norm_vars=norm_vars)
    ref = oracle.cmvn_apply(mfcc.data, ref_stats, norm_vars=norm_vars)
    assert np.allclose(out.data, ref, rtol=1e-5, atol=1e-5)
    assert np.abs(out.data.mean(0)).max() < 1e-5          # test_cmvn.py:40-60
    if norm_vars:
        assert np.abs(out.data.var(0) - 1).max() < 1e-4
    back = proc.process(out, norm_vars=norm_vars, reverse=True)
    assert np.abs(back.data - mfcc.data).max() < 1e-3
    assert np.array_equal(mfcc.data, before)              # input not mutated
    assert np.array_equal(out.properties['cmvn']['stats'], proc.stats)
    # stats doubling, weights -> count, skip_dims
    proc.accumulate(mfcc)
    assert np.allclose(proc.stats, 2 * ref_stats, rtol=1e-12)
    w = np.zeros(140)
    w[:70] = 1
    p2 = CmvnPostProcessor(13)
    p2.accumulate(mfcc, weights=w)
    assert p2.count == 70
    assert np.allclose(
        p2.stats, oracle.cmvn_accumulate(mfcc.data, w), rtol=1e-12)
    skipped = p2.process(mfcc, skip_dims=[0, 5])
    assert np.array_equal(skipped.data[:, [0, 5]], mfcc.data[:, [0, 5]])
    with pytest.raises(ValueError):
        p2.process(mfcc, skip_dims=[13])
    with pytest.raises(ValueError):
        p2.accumulate(mfcc, weights=np.zeros((140, 1)))


def test_apply_cmvn_collection(mfcc):
    other = Features(mfcc.data[::-1].copy() * 1.5 + 2, mfcc.times.copy())
    coll = FeaturesCollection(a=mfcc, b=other)
    by_coll = apply_cmvn(coll, by_collection=True)
    stacked = np.vstack([by_coll['a'].data, by_coll['b'].data])
    assert np.abs(stacked.mean(0)).max() < 1e-5
    by_item = apply_cmvn(coll, by_collection=False)
    for f in by_item.values():
        assert np.abs(f.data.mean(0)).max() < 1e-5


def test_sliding_window_cmn(golden, mfcc):
    data, manifest = golden
    base = Features(data['mfcc_default'], mfcc.times)
    for name, entry in manifest.items():
        if entry['kind'] != 'sliding_window_cmn':
            continue
        kw = entry['kwargs']
        out = SlidingWindowCmvnPostProcessor(
            center=kw['center'], cmn_window=kw['cmn_window'],
            min_window=kw['min_cmn_window'],
            normalize_variance=kw['norm_vars']).process(base)
        scale_close(out.data, data[name], tol=1e-4)
        ref = oracle.sliding_window_cmn(
            base.data, center=kw['center'], cmn_window=kw['cmn_window'],
            min_window=kw['min_cmn_window'], normalize_variance=kw['norm_vars'])
        assert np.allclose(out.data, ref, atol=1e-5)


@pytest.mark.parametrize('kwargs', [
    {}, {'frames_context': 2}, {'energy_threshold': 10, 'energy_mean_scale': 0},
    {'frames_context': 5, 'proportion_threshold': 0.3}])
def test_vad(mfcc, kwargs):
    out = VadPostProcessor(**kwargs).process(mfcc)
    ref = oracle.vad(mfcc.data, **kwargs)
    assert out.dtype == np.uint8 and out.shape == (140, 1)
    assert np.array_equal(out.data, ref)
    assert 0 < out.data.sum() < 140
