"""Generates tests/golden/plp_reference_shim.npz (run in the build container).

Runs the UNMODIFIED reference module ``shennong.processor.plp`` from
/root/reference (PLP and RASTA-PLP are pure Python there, plp.py:64-626) on the
reference's own test input, with the absent pykaldi package replaced by the
numpy stand-in of tests/golden/pykaldi_shim (mel banks from
torchaudio.compliance.kaldi, the other primitives restated from Kaldi in
float32; see its README).  What these vectors pin is therefore the
reference's own frame loop / ExtractWindow / ProcessWindow / RASTA state
machine / LPC -> cepstrum / lifter / energy / HTK ordering, executed as is.

tests/test_oracle_golden.py checks the CPU oracle against this file and the
GPU tests check the CUDA path against it (the file travels to the GPU box,
/root/reference does not).

Usage:  python tests/golden/make_plp_golden.py
"""

import json
import os
import sys
import warnings

import numpy as np
import scipy.io.wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'pykaldi_shim'))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')

import types  # noqa: E402

from shennong import Audio  # noqa: E402

# `shennong/processor/__init__.py` imports every processor (bottleneck, CREPE,
# UBM/VTLN training ...) and with them pykaldi modules far off this path: register
# an empty package in its place and import the two modules the recipe needs,
# unmodified, from the reference tree.
_pkg = types.ModuleType('shennong.processor')
_pkg.__path__ = ['/root/reference/shennong/processor']
sys.modules['shennong.processor'] = _pkg
from shennong.processor.plp import PlpProcessor  # noqa: E402
from shennong.processor.energy import EnergyProcessor  # noqa: E402

WAV = '/root/reference/test/data/test.wav'

# (name, constructor kwargs, vtln_warp); dither is always 0 (deterministic)
CASES = [
    ('plp_default', {}, 1.0),
    ('plp_c0', {'use_energy': False}, 1.0),
    ('plp_htk', {'htk_compat': True}, 1.0),
    ('plp_not_raw_energy', {'raw_energy': False}, 1.0),
    ('plp_energy_floor', {'energy_floor': 1.0e7}, 1.0),
    ('plp_lpc8_ceps9', {'lpc_order': 8, 'num_ceps': 9}, 1.0),
    ('plp_lpc14_ceps10', {'lpc_order': 14, 'num_ceps': 10}, 1.0),
    ('plp_nolifter_scale', {'cepstral_lifter': 0.0, 'cepstral_scale': 2.0}, 1.0),
    ('plp_compress_half', {'compress_factor': 0.5}, 1.0),
    ('plp_30bins_lowhigh', {'num_bins': 30, 'low_freq': 100, 'high_freq': -400}, 1.0),
    ('plp_hamming_nodc', {'window_type': 'hamming', 'remove_dc_offset': False}, 1.0),
    ('plp_nosnip', {'snip_edges': False}, 1.0),
    ('plp_vtln_1.1', {}, 1.1),
    ('plp_vtln_0.9', {}, 0.9),
    ('plp_shift20_len50', {'frame_shift': 0.02, 'frame_length': 0.05}, 1.0),
    ('rasta_default', {'rasta': True}, 1.0),
    ('rasta_c0_htk', {'rasta': True, 'use_energy': False, 'htk_compat': True}, 1.0),
    ('rasta_vtln_1.1', {'rasta': True}, 1.1),
]


# EnergyProcessor is a Python frame loop as well (energy.py:153-186)
ENERGY_CASES = [
    ('energy_default', {}),
    ('energy_sqrt', {'compression': 'sqrt'}),
    ('energy_off_not_raw', {'compression': 'off', 'raw_energy': False}),
    ('energy_hamming_not_raw', {'raw_energy': False, 'window_type': 'hamming'}),
    ('energy_nosnip', {'snip_edges': False}),
    ('energy_shift20_len50_nodc', {'frame_shift': 0.02, 'frame_length': 0.05, 'remove_dc_offset': False}),
]


def main():
    rate, pcm = scipy.io.wavfile.read(WAV)
    assert rate == 16000 and pcm.dtype == np.int16 and pcm.ndim == 1
    audio = Audio(pcm, rate, validate=False)
    out = {'pcm': pcm}
    meta = {}
    for name, kwargs, warp in CASES:
        proc = PlpProcessor(sample_rate=rate, dither=0.0, **kwargs)
        feats = proc.process(audio, vtln_warp=warp)
        out[name] = np.asarray(feats.data, dtype=np.float32)
        out[name + '.times'] = np.asarray(feats.times, dtype=np.float64)
        meta[name] = {'kind': 'plp', 'kwargs': kwargs, 'vtln_warp': warp, 'shape': list(feats.shape)}
        print('%-24s %s' % (name, feats.shape))
    for name, kwargs in ENERGY_CASES:
        proc = EnergyProcessor(sample_rate=rate, dither=0.0, **kwargs)
        feats = proc.process(audio)
        assert feats.data.dtype == np.float64
        out[name] = np.asarray(feats.data, dtype=np.float64)
        out[name + '.times'] = np.asarray(feats.times, dtype=np.float64)
        meta[name] = {'kind': 'energy', 'kwargs': kwargs, 'shape': list(feats.shape)}
        print('%-24s %s' % (name, feats.shape))
    out['meta'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, 'plp_reference_shim.npz'), **out)


if __name__ == '__main__':
    main()
