"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/snb.h declares; its host-side helpers agree with the oracle."""

import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT
from shennong_b200 import _lib


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'snb.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(snb_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported_and_bound():
    symbols = header_symbols()
    assert len(symbols) >= 35
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in symbols:
        assert hasattr(handle, name), f'{name} missing from libsnb.so'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
    assert sorted(_lib.SIGNATURES) == symbols
    assert _lib.lib().snb_version() == 100


@pytest.mark.parametrize('kw', [
    {}, {'snip_edges': False}, {'frame_shift': 0.02, 'frame_length': 0.05},
    {'sample_rate': 8000}, {'sample_rate': 44100},
    {'round_to_power_of_two': False}])
def test_framing_matches_oracle(kw):
    L = _lib.lib()
    fo = oracle.frame_opts(**kw)
    mine = _lib.make_frame_opts(
        kw.get('sample_rate', 16000), kw.get('frame_shift', 0.01) * 1000.0,
        kw.get('frame_length', 0.025) * 1000.0, 0.0, 0.97, True, 'povey',
        kw.get('round_to_power_of_two', True), 0.42,
        kw.get('snip_edges', True))
    O = oracle.lib()
    for fun in ('window_size', 'window_shift', 'padded_window_size'):
        assert (getattr(L, 'snb_' + fun)(_lib.ref(mine))
                == getattr(O, 'orc_' + fun)(ctypes.byref(fo)))
    for n in [0, 1, 399, 400, 401, 559, 560, 22713, 160000, 1234567]:
        assert (L.snb_num_frames(n, _lib.ref(mine))
                == O.orc_num_frames(n, ctypes.byref(fo)))
    from shennong_b200 import engine
    ns = np.array([0, 1, 399, 400, 401, 559, 560, 22713, 160000, 1234567])
    assert np.array_equal(
        engine.num_frames_array(mine, ns),
        [L.snb_num_frames(int(n), _lib.ref(mine)) for n in ns])
    for f in [0, 1, 7, 139]:
        assert (L.snb_first_sample_of_frame(f, _lib.ref(mine))
                == O.orc_first_sample_of_frame(f, ctypes.byref(fo)))


@pytest.mark.parametrize('wtype', sorted(_lib.WINDOW_TYPES))
def test_window_matches_oracle(wtype):
    from shennong_b200.window import window
    for length in (5, 400, 441):
        assert np.array_equal(
            window(length, wtype, 0.4), oracle.window(length, wtype, 0.4))


@pytest.mark.parametrize('kw,warp', [
    ({}, 1.0), ({'num_bins': 40}, 1.0), ({}, 1.1), ({}, 0.85),
    ({'num_bins': 30, 'low_freq': 100, 'high_freq': -400}, 1.0),
    ({'sample_rate': 8000}, 1.0), ({'round_to_power_of_two': False}, 0.9)])
def test_mel_banks_match_oracle(kw, warp):
    ref_w, ref_c = oracle.mel_banks(vtln_warp=warp, **kw)
    fo = _lib.make_frame_opts(
        kw.get('sample_rate', 16000), 10.0, 25.0, 0.0, 0.97, True, 'povey',
        kw.get('round_to_power_of_two', True), 0.42, True)
    mo = _lib.MelOpts(kw.get('num_bins', 23), kw.get('low_freq', 20),
                      kw.get('high_freq', 0), 100, -500)
    w = np.zeros_like(ref_w)
    c = np.zeros_like(ref_c)
    _lib.check(_lib.lib().snb_mel_banks_host(
        _lib.ref(fo), _lib.ref(mo), np.float32(warp), _lib.np_ptr(w),
        _lib.np_ptr(c)))
    assert np.array_equal(w, ref_w) and np.array_equal(c, ref_c)


def test_mel_options_kaldi_rejects():
    fo = _lib.make_frame_opts(16000, 10.0, 25.0, 0, 0.97, True, 'povey',
                              True, 0.42, True)
    w, c = np.zeros((23, 256), np.float32), np.zeros(23, np.float32)
    for mo in (_lib.MelOpts(2, 20, 0, 100, -500),
               _lib.MelOpts(23, 9000, 0, 100, -500),
               _lib.MelOpts(23, 20, 10, 100, -500)):
        with pytest.raises(RuntimeError):
            _lib.check(_lib.lib().snb_mel_banks_host(
                _lib.ref(fo), _lib.ref(mo), 1.0, _lib.np_ptr(w),
                _lib.np_ptr(c)))


def test_pitch_frame_count_matches_oracle():
    po_mine = _lib.PitchOpts(16000, 10.0, 25.0, 0.0, 50, 400, 10, 0.1, 1000,
                             4000, 0.005, 7000, 1, 5, 1)
    po = oracle.pitch_opts()
    for n in [0, 100, 399, 400, 1000, 22713, 160000, 160001, 99999]:
        assert (_lib.lib().snb_pitch_num_frames(n, _lib.ref(po_mine))
                == oracle.lib().orc_pitch_num_frames(n, ctypes.byref(po)))
    assert _lib.lib().snb_pitch_num_frames(22713, _lib.ref(po_mine)) == 140
    assert _lib.lib().snb_pitch_num_frames(160000, _lib.ref(po_mine)) == 998


def test_compute_fails_loudly_without_gpu(audio):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from shennong_b200.processor import MfccProcessor
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        MfccProcessor().process(audio)
