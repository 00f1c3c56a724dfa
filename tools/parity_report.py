#!/usr/bin/env python
"""Parity evidence for profiles/: the CUDA path (through the C ABI) against every
committed golden vector and against the CPU oracle, as numbers rather than
pass/fail.

    python tools/parity_report.py > profiles/rNN_parity_report.txt     (needs a B200)

Per case: max|a-b| / max|ref| (the gate of the tests, 1e-4), and the
element-wise relative error over the elements with |ref| >= 1e-3 max|ref|
(median / 99th percentile / max).  Frame times are compared bit for bit.
"""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def stats(out, ref):
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(out - ref)
    big = np.abs(ref) >= 1e-3 * scale
    rel = err[big] / np.abs(ref[big])
    return (err.max() / scale, np.median(rel), np.percentile(rel, 99), rel.max())


def main():
    import oracle
    from conftest import synth_utterance
    from shennong_b200 import Audio
    from shennong_b200.processor import (
        EnergyProcessor, FilterbankProcessor, MfccProcessor, PlpProcessor,
        SpectrogramProcessor)
    procs = {'mfcc': MfccProcessor, 'filterbank': FilterbankProcessor,
             'spectrogram': SpectrogramProcessor, 'plp': PlpProcessor,
             'energy': EnergyProcessor}
    print('# case, frames x dims, max|a-b|/max|ref|, elementwise rel err '
          '(|ref| >= 1e-3 max|ref|): median, p99, max, times bit-equal')
    worst = 0.0

    def row(name, feats, ref, times=None):
        nonlocal worst
        s = stats(feats.data, ref)
        worst = max(worst, s[0])
        teq = '' if times is None else str(bool(np.array_equal(feats.times, times)))
        print(f'{name:34s} {feats.shape[0]:5d} x {feats.shape[1]:<3d} '
              f'{s[0]:.2e}  {s[1]:.2e} {s[2]:.2e} {s[3]:.2e}  {teq}')

    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'kaldi_compliance.npz'))
    manifest = json.loads(bytes(g['manifest']).decode())
    pcm = g['pcm']
    print('## tests/golden/kaldi_compliance.npz (torchaudio.compliance.kaldi on test.wav)')
    for name, entry in manifest.items():
        if entry['kind'] == 'sliding_window_cmn':
            continue
        kw = dict(entry['kwargs'])
        warp = kw.pop('vtln_warp', None)
        rate = kw.pop('sample_rate', 16000)
        proc = procs[entry['kind']](sample_rate=rate, dither=0, **kw)
        audio = Audio(pcm, rate)
        feats = proc.process(audio) if warp is None else proc.process(audio, vtln_warp=warp)
        row(name, feats, g[name])
    p = np.load(os.path.join(ROOT, 'tests', 'golden', 'plp_reference_shim.npz'))
    meta = json.loads(str(p['meta']))
    print('## tests/golden/plp_reference_shim.npz (the reference\'s plp.py / energy.py over the pykaldi shim)')
    for name, entry in meta.items():
        kw = dict(entry['kwargs'])
        proc = procs[entry['kind']](dither=0, **kw)
        warp = entry.get('vtln_warp', 1.0)
        audio = Audio(pcm, 16000)
        feats = proc.process(audio) if warp == 1.0 else proc.process(audio, vtln_warp=warp)
        row(name, feats, p[name], p[name + '.times'])
    print('## CPU oracle on a 10 s synthetic utterance (998 frames)')
    sig = synth_utterance(3, 160000)
    for kind in ('mfcc', 'filterbank', 'spectrogram', 'plp'):
        feats = procs[kind](dither=0).process(Audio(sig, 16000))
        row(f'{kind} (oracle)', feats, oracle.features(kind, sig))
    print(f'# worst max|a-b|/max|ref| = {worst:.2e} (gate 1e-4)')
    pitch_rows(oracle, pcm, synth_utterance)


def pitch_rows(oracle, pcm, synth_utterance):
    """Kaldi pitch: index work (the Viterbi state of every frame) must be
    bit-exact against the oracle; known answers the oracle cannot fake"""
    from conftest import numpy_process_pitch
    from shennong_b200 import Audio, Features
    from shennong_b200.processor import (
        KaldiPitchPostProcessor, KaldiPitchProcessor)
    print('## Kaldi pitch against the CPU oracle: frames, frames on the same '
          'Viterbi state, NCCF column bit-equal, max|NCCF diff|')
    cases = [('test.wav defaults', pcm, {}),
             ('test.wav min_f0=60 max_f0=350', pcm, {'min_f0': 60, 'max_f0': 350}),
             ('test.wav shift 20 ms length 50 ms', pcm,
              {'frame_shift': 0.02, 'frame_length': 0.05}),
             ('test.wav penalty_factor 0.3', pcm, {'penalty_factor': 0.3}),
             ('test.wav delta_pitch 0.002 (1040 states)', pcm, {'delta_pitch': 0.002}),
             ('test.wav upsample_filter_width 7', pcm, {'upsample_filter_width': 7})]
    cases += [(f'synthetic 10 s utterance {u}', synth_utterance(u, 160000), {})
              for u in (3, 11, 42)]
    all_same = True
    for name, sig, kw in cases:
        out = KaldiPitchProcessor(**kw).process(Audio(sig, 16000)).data
        ref = oracle.pitch(sig, **kw)
        same = int((out[:, 1] == ref[:, 1]).sum())
        all_same = all_same and same == len(ref) and np.array_equal(out[:, 0], ref[:, 0])
        print(f'{name:44s} {len(ref):5d} {same:5d}  '
              f'{bool(np.array_equal(out[:, 0], ref[:, 0]))}  '
              f'{np.abs(out[:, 0] - ref[:, 0]).max():.1e}')
    print(f'# state sequence and NCCF bit-exact on every case: {all_same}')
    print('## known answers: 5-harmonic tone of fundamental f0 (3 s), median '
          '|f0_est / f0 - 1| over the frames (lag grid step 0.5 %)')
    t = np.arange(48000) / 16000.0
    for f0 in (60, 85, 120, 170, 240, 350):
        x = sum(3000.0 / h * np.sin(2 * np.pi * h * f0 * t + 0.3 * h) for h in range(1, 6))
        sig = np.round(x).astype(np.int16)
        out = KaldiPitchProcessor(min_f0=50, max_f0=400).process(Audio(sig, 16000)).data
        rel = np.abs(out[:, 1] / f0 - 1.0)
        print(f'f0 = {f0:3d} Hz   median {np.median(rel):.2e}   p90 {np.percentile(rel, 90):.2e}'
              f'   median NCCF {np.median(out[:, 0]):.3f}')
    print('## post-processing against a float64 numpy evaluation of the published '
          'formulas (pitch_crepe.py:246-253 for POV): max abs error per column '
          '(pov feature, normalised log-pitch, delta, raw log-pitch)')
    raw = KaldiPitchProcessor().process(Audio(pcm, 16000))
    post = KaldiPitchPostProcessor(delta_pitch_noise_stddev=0, add_raw_log_pitch=True).process(raw)
    want = numpy_process_pitch(raw.data)
    print('test.wav  ' + '  '.join(f'{e:.1e}' for e in np.abs(post.data - want).max(axis=0)))
    ref = oracle.process_pitch(raw.data, delta_pitch_noise_stddev=0, add_raw_log_pitch=True)
    print('against the oracle: ' + '  '.join(f'{e:.1e}' for e in np.abs(post.data - ref).max(axis=0)))


if __name__ == '__main__':
    main()
