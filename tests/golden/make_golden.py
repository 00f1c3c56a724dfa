"""Generates tests/golden/kaldi_compliance.npz (run in the build container).

The reference's numeric backend (pykaldi/Kaldi) is absent from the container
and from /root/reference, and the reference's tests hold no golden numeric
vectors for this path.  The closest executable, independent
Kaldi-compliance implementation available offline is
``torchaudio.compliance.kaldi`` (torchaudio 2.11): this script runs it on the
reference's own correctness input (test/data/test.wav, 16 kHz int16 mono,
22 713 samples -- the input BASELINE.json configs[0] names) for a grid of
option sets and stores inputs + outputs.  tests/test_oracle_golden.py checks
the CPU oracle against these vectors; the GPU tests check the CUDA path
against the same file (it travels to the GPU box, /root/reference does not).

Usage:  python tests/golden/make_golden.py
"""

import json
import os

import numpy as np
import scipy.io.wavfile
import torch
import torchaudio
import torchaudio.compliance.kaldi as K

HERE = os.path.dirname(os.path.abspath(__file__))
WAV = '/root/reference/test/data/test.wav'

# (name, kind, shennong-style kwargs); translated to torchaudio kwargs below
CASES = [
    ('mfcc_default', 'mfcc', {}),
    ('mfcc_c0', 'mfcc', {'use_energy': False}),
    ('mfcc_htk', 'mfcc', {'htk_compat': True, 'use_energy': False}),
    ('mfcc_nolifter_20ceps', 'mfcc', {'cepstral_lifter': 0.0, 'num_ceps': 20}),
    ('mfcc_not_raw_energy', 'mfcc', {'raw_energy': False}),
    ('mfcc_energy_floor', 'mfcc', {'energy_floor': 1.0e7}),
    ('mfcc_hamming', 'mfcc', {'window_type': 'hamming'}),
    ('mfcc_nosnip', 'mfcc', {'snip_edges': False}),
    ('mfcc_vtln_1.1', 'mfcc', {'vtln_warp': 1.1}),
    ('mfcc_vtln_0.9', 'mfcc', {'vtln_warp': 0.9}),
    ('mfcc_8k', 'mfcc', {'sample_rate': 8000}),
    ('mfcc_shift20_len50', 'mfcc', {'frame_shift': 0.02, 'frame_length': 0.05}),
    ('mfcc_nopow2', 'mfcc', {'round_to_power_of_two': False}),
    ('fbank_40', 'filterbank', {'num_bins': 40}),
    ('fbank_23_energy', 'filterbank', {'use_energy': True}),
    ('fbank_23_energy_htk', 'filterbank', {'use_energy': True,
                                           'htk_compat': True}),
    ('fbank_linear', 'filterbank', {'use_log_fbank': False}),
    ('fbank_magnitude', 'filterbank', {'use_power': False}),
    ('fbank_lowhigh', 'filterbank', {'low_freq': 100, 'high_freq': -400,
                                     'num_bins': 30}),
    ('fbank_nodc_nopre_rect', 'filterbank', {
        'remove_dc_offset': False, 'preemph_coeff': 0.0,
        'window_type': 'rectangular'}),
    ('fbank_hanning', 'filterbank', {'window_type': 'hanning'}),
    ('fbank_blackman', 'filterbank', {'window_type': 'blackman',
                                      'blackman_coeff': 0.4}),
    ('spectrogram_default', 'spectrogram', {}),
    ('spectrogram_not_raw', 'spectrogram', {'raw_energy': False}),
]

RENAME = {'sample_rate': 'sample_frequency', 'num_bins': 'num_mel_bins',
          'preemph_coeff': 'preemphasis_coefficient'}


def run_case(wave, kind, kwargs):
    kw = {'dither': 0.0, 'energy_floor': 0.0}
    if kind == 'mfcc':
        kw.update(use_energy=True, num_mel_bins=23, num_ceps=13)
    for k, v in kwargs.items():
        if k in ('frame_shift', 'frame_length'):
            kw[k] = v * 1000.0
        else:
            kw[RENAME.get(k, k)] = v
    kw.setdefault('sample_frequency', 16000.0)
    kw['sample_frequency'] = float(kw['sample_frequency'])
    fun = {'mfcc': K.mfcc, 'filterbank': K.fbank,
           'spectrogram': K.spectrogram}[kind]
    return fun(wave, **kw).numpy()


def main():
    rate, pcm = scipy.io.wavfile.read(WAV)
    assert rate == 16000 and pcm.dtype == np.int16 and pcm.shape == (22713,)
    wave = torch.from_numpy(pcm.astype(np.float32))[None]
    out = {'pcm': pcm, 'sample_rate': np.int64(rate)}
    manifest = {}
    for name, kind, kwargs in CASES:
        out[name] = run_case(wave, kind, kwargs).astype(np.float32)
        manifest[name] = {'kind': kind, 'kwargs': kwargs,
                          'shape': list(out[name].shape)}
        print(name, out[name].shape)

    # sliding-window CMN (torchaudio.functional follows Kaldi's
    # SlidingWindowCmn); inputs: the default MFCCs
    base = torch.from_numpy(out['mfcc_default'])
    for name, kw in [
            ('swcmn_center_600', dict(cmn_window=600, min_cmn_window=100,
                                      center=True, norm_vars=False)),
            ('swcmn_center_50_var', dict(cmn_window=50, min_cmn_window=10,
                                         center=True, norm_vars=True)),
            ('swcmn_left_60', dict(cmn_window=60, min_cmn_window=20,
                                   center=False, norm_vars=False))]:
        out[name] = torchaudio.functional.sliding_window_cmn(
            base, **kw).numpy().astype(np.float32)
        manifest[name] = {'kind': 'sliding_window_cmn', 'kwargs': kw,
                          'shape': list(out[name].shape)}
        print(name, out[name].shape)

    out['manifest'] = np.frombuffer(
        json.dumps(manifest).encode(), dtype=np.uint8)
    out['versions'] = np.frombuffer(json.dumps({
        'torchaudio': torchaudio.__version__,
        'torch': torch.__version__}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, 'kaldi_compliance.npz'), **out)


if __name__ == '__main__':
    main()
