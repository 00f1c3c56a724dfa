"""GPU parity of the fused feature kernels against the CPU oracle and the
committed golden vectors (all calls go through the C ABI)."""

import numpy as np
import pytest

import oracle
from conftest import scale_close, synth_utterance
from shennong_b200 import Audio
from shennong_b200.processor import (
    EnergyProcessor, FilterbankProcessor, MfccProcessor, PlpProcessor,
    SpectrogramProcessor)

pytestmark = pytest.mark.gpu

PROCESSORS = {'mfcc': MfccProcessor, 'filterbank': FilterbankProcessor,
              'spectrogram': SpectrogramProcessor, 'plp': PlpProcessor}


def run(kind, pcm, rate=16000, vtln_warp=None, **kwargs):
    kwargs.setdefault('dither', 0)
    proc = PROCESSORS[kind](sample_rate=rate, **kwargs)
    audio = Audio(pcm, rate)
    if vtln_warp is None:
        return proc.process(audio)
    return proc.process(audio, vtln_warp=vtln_warp)


def test_all_golden_vectors(golden):
    """Every option set of the golden file (fast path, generic path, VTLN,
    non-pow2 FFT, 8 kHz, snip_edges=False ...)"""
    data, manifest = golden
    pcm = data['pcm']
    for name, entry in manifest.items():
        if entry['kind'] == 'sliding_window_cmn':
            continue
        kwargs = dict(entry['kwargs'])
        warp = kwargs.pop('vtln_warp', None)
        rate = kwargs.pop('sample_rate', 16000)
        feats = run(entry['kind'], pcm, rate=rate, vtln_warp=warp, **kwargs)
        try:
            scale_close(feats.data, data[name], tol=1e-4)
        except AssertionError as err:
            raise AssertionError(f'{name}: {err}') from None


def test_plp_vs_reference_plp_py(golden_plp):
    """PLP / RASTA-PLP (fused kernel, RASTA workspace path, VTLN, generic FFT
    path) against the outputs of the reference's own plp.py, times included"""
    data, meta = golden_plp
    pcm = data['pcm']
    for name, entry in meta.items():
        if entry['kind'] == 'energy':
            feats = EnergyProcessor(dither=0, **entry['kwargs']).process(
                Audio(pcm, 16000))
            assert feats.dtype == np.float64
        else:
            warp = entry['vtln_warp']
            feats = run('plp', pcm, vtln_warp=None if warp == 1.0 else warp,
                        **entry['kwargs'])
        try:
            # (energy: float64 sums of float32 products on both sides; the
            # compressed value agrees far below the gate)
            scale_close(feats.data, data[name], tol=1e-4)
        except AssertionError as err:
            raise AssertionError(f'{name}: {err}') from None
        assert np.array_equal(feats.times, data[name + '.times']), name


@pytest.mark.parametrize('kind,kwargs', [
    ('mfcc', {}),
    ('mfcc', {'use_energy': False, 'htk_compat': True}),
    ('mfcc', {'raw_energy': False, 'energy_floor': 3.0e6}),
    ('mfcc', {'num_bins': 40, 'num_ceps': 40, 'cepstral_lifter': 0}),
    ('mfcc', {'remove_dc_offset': False, 'preemph_coeff': 0.0,
              'window_type': 'hamming'}),
    ('mfcc', {'frame_length': 0.02, 'frame_shift': 0.005}),
    ('mfcc', {'frame_length': 0.032, 'frame_shift': 0.0125,
              'snip_edges': False}),
    ('filterbank', {'num_bins': 40}),
    ('filterbank', {'num_bins': 80, 'use_energy': True}),
    ('filterbank', {'use_power': False, 'use_log_fbank': False}),
    ('spectrogram', {}),
    ('spectrogram', {'raw_energy': False, 'window_type': 'blackman'}),
    ('plp', {}),
    ('plp', {'use_energy': False, 'htk_compat': True}),
    ('plp', {'lpc_order': 10, 'num_ceps': 8, 'cepstral_scale': 2.0,
             'compress_factor': 0.5, 'num_bins': 30}),
    ('plp', {'snip_edges': False, 'raw_energy': False}),
    # odd window length (321 samples) and odd shift (161): the element-wise
    # load path and the ragged last register of the fused kernel
    ('mfcc', {'frame_length': 0.0200625, 'frame_shift': 0.0100625}),
    ('filterbank', {'frame_length': 0.0200625, 'remove_dc_offset': False}),
    # 100 ms shift: tiles of 8 frames (half of the lane groups idle)
    ('mfcc', {'frame_shift': 0.1}),
    ('spectrogram', {'frame_shift': 0.05, 'snip_edges': False}),
])
def test_fast_path_vs_oracle(pcm, kind, kwargs):
    feats = run(kind, pcm, **kwargs)
    ref = oracle.features(kind, pcm, **kwargs)
    assert feats.dtype == np.float32
    scale_close(feats.data, ref, tol=1e-4)
    assert np.isfinite(feats.data).all()


@pytest.mark.parametrize('kind,rate,kwargs', [
    ('mfcc', 8000, {}),                                   # N = 256
    ('mfcc', 44100, {}),                                  # W = 1102, N = 2048
    ('filterbank', 16000, {'frame_length': 0.05}),        # N = 1024
    ('mfcc', 16000, {'round_to_power_of_two': False}),    # direct DFT, N = 400
    ('plp', 8000, {'num_bins': 15}),
    ('spectrogram', 16000, {'frame_length': 0.01}),       # W = 160, N = 256
    ('plp', 16000, {'lpc_order': 20, 'num_ceps': 21, 'num_bins': 30}),
])
def test_generic_path_vs_oracle(pcm, kind, rate, kwargs):
    feats = run(kind, pcm, rate=rate, **kwargs)
    ref = oracle.features(kind, pcm, sample_rate=rate, **kwargs)
    scale_close(feats.data, ref, tol=1e-4)


@pytest.mark.parametrize('warp', [0.85, 0.95, 1.05, 1.25])
def test_vtln_warp(pcm, warp):
    for kind in ('mfcc', 'filterbank', 'plp'):
        feats = run(kind, pcm, vtln_warp=warp)
        ref = oracle.features(kind, pcm, vtln_warp=warp)
        scale_close(feats.data, ref, tol=1e-4)
        assert feats.properties[PROCESSORS[kind]().name]['vtln_warp'] == warp


def test_frame_counts_and_times_bit_exact(pcm):
    for kwargs, nframes in [({}, 140), ({'frame_shift': 0.02}, 70),
                            ({'frame_shift': 0.02, 'frame_length': 0.05}, 69),
                            ({'snip_edges': False}, 142)]:
        feats = run('mfcc', pcm, **kwargs)
        assert feats.shape == (nframes, 13)
        shift = np.float32(np.float32(kwargs.get('frame_shift', 0.01) * 1000.0)
                           / 1000.0)
        length = np.float32(
            np.float32(kwargs.get('frame_length', 0.025) * 1000.0) / 1000.0)
        start = np.arange(nframes) * shift
        assert feats.times.dtype == np.float64
        assert np.array_equal(feats.times, np.vstack((start, start + length)).T)


def test_dtype_independence(pcm):
    """int16 / int32 / float32 / float64 inputs give the same features
    (test/processor/test_mfcc.py:145-173)"""
    ref = run('mfcc', pcm).data
    a32 = Audio(pcm.astype(np.int32) * 2**15, 16000)
    f32 = Audio((pcm / 2**15).astype(np.float32), 16000)
    f64 = Audio(pcm / 2**15, 16000)
    proc = MfccProcessor(dither=0)
    for audio in (a32, f32, f64):
        assert np.array_equal(proc.process(audio).data, ref)


def test_energy_processor(pcm):
    audio = Audio(pcm, 16000)
    for kwargs in ({}, {'compression': 'off'}, {'compression': 'sqrt'},
                   {'raw_energy': False, 'window_type': 'hanning'},
                   {'frame_shift': 0.02, 'frame_length': 0.05}):
        energy = EnergyProcessor(dither=0, **kwargs).process(audio)
        ref = oracle.features('energy', pcm, **kwargs)
        assert energy.dtype == np.float64 and energy.shape == ref.shape
        assert np.allclose(energy.data, ref, rtol=2e-6)
    # equals column 0 of MFCC and PLP (test/processor/test_energy.py:36-44)
    energy = EnergyProcessor(dither=0).process(audio)
    assert np.allclose(energy.data[:, 0], run('mfcc', pcm).data[:, 0],
                       rtol=1e-5)
    assert np.allclose(energy.data[:, 0], run('plp', pcm).data[:, 0],
                       rtol=1e-5)
    # float audio keeps its [-1, 1] scale (energy.py:158 has no int16 cast)
    faudio = Audio((pcm / 2**15).astype(np.float32), 16000)
    fen = EnergyProcessor(dither=0).process(faudio)
    fref = oracle.features('energy', (pcm / 2**15).astype(np.float32))
    assert np.allclose(fen.data, fref, rtol=1e-5, atol=1e-5)


def test_invalid_options_raise_like_kaldi(audio):
    # test/processor/test_mfcc.py:69-97, test_filterbank.py:41-50
    for num_ceps in (0, 25):
        with pytest.raises(RuntimeError):
            MfccProcessor(num_ceps=num_ceps).process(audio)
    for num_bins in (0, 1, 2):
        with pytest.raises(RuntimeError):
            FilterbankProcessor(num_bins=num_bins).process(audio)
        with pytest.raises(RuntimeError):
            MfccProcessor(num_bins=num_bins, num_ceps=1).process(audio)
    with pytest.raises(ValueError):
        MfccProcessor(sample_rate=8000).process(audio)
    with pytest.raises(ValueError):
        stereo = Audio(np.zeros((1000, 2), np.int16), 16000)
        MfccProcessor().process(stereo)


@pytest.mark.parametrize('kwargs', [
    {}, {'use_energy': False}, {'raw_energy': False, 'num_bins': 30},
    {'frame_length': 0.05}, {'htk_compat': True, 'cepstral_lifter': 0}])
def test_rasta_plp(pcm, kwargs):
    """RASTA-PLP (plp.py:64-146): first 4 frames see unit mel energies, then
    the float64 IIR; fast and generic paths"""
    feats = run('plp', pcm, rasta=True, **kwargs)
    ref = oracle.features('plp', pcm, rasta=True, **kwargs)
    scale_close(feats.data, ref, tol=1e-4)
    plain = run('plp', pcm, **kwargs)
    assert not np.allclose(plain.data[10:, 1:], feats.data[10:, 1:], atol=1e-2)
    # short utterances: fewer than 4 frames never leave the priming phase
    short = synth_utterance(3, 800)
    scale_close(run('plp', short, rasta=True).data,
                oracle.features('plp', short, rasta=True), tol=1e-4)
    long = synth_utterance(5, 160000)
    scale_close(run('plp', long, rasta=True).data,
                oracle.features('plp', long, rasta=True), tol=1e-4)


def test_edge_lengths():
    """empty / shorter than a frame / exactly one frame / ragged tails"""
    rng = np.random.default_rng(1)
    proc = MfccProcessor(dither=0)
    for n in (0, 10, 399):
        feats = proc.process(Audio(np.zeros(n, np.int16), 16000))
        assert feats.shape[0] == 0
    for n in (400, 401, 559, 560, 561, 5517, 12345):
        pcm = rng.integers(-20000, 20000, n).astype(np.int16)
        feats = proc.process(Audio(pcm, 16000))
        ref = oracle.features('mfcc', pcm)
        assert feats.shape == ref.shape
        scale_close(feats.data, ref, tol=1e-4)
    # all-zero signal: log floors (FLT_EPSILON) must match, no NaN
    zeros = np.zeros(3000, np.int16)
    for kind in ('mfcc', 'filterbank', 'spectrogram'):
        feats = run(kind, zeros)
        # (the DCT of a constant vector is rounding noise around 0)
        assert np.allclose(feats.data, oracle.features(kind, zeros),
                           rtol=0, atol=1e-4)
        if kind != 'mfcc':
            assert np.array_equal(feats.data, oracle.features(kind, zeros))
    # full-scale square wave (maximum magnitudes)
    square = (np.sign(np.sin(np.arange(8000) * 0.05)) * 32767).astype(np.int16)
    scale_close(run('mfcc', square).data, oracle.features('mfcc', square))


def test_ragged_batch_matches_single_calls(tmp_path):
    """process_all on utterances of different lengths == per-utterance
    process (one launch, tiles never cross utterances)"""
    import scipy.io.wavfile
    from shennong_b200 import Utterances
    lengths = [160000, 401, 9000, 31999, 16000, 400, 123457]
    utts = []
    for i, n in enumerate(lengths):
        path = tmp_path / f'u{i}.wav'
        scipy.io.wavfile.write(path, 16000, synth_utterance(i, n))
        utts.append((f'utt{i}', str(path)))
    utterances = Utterances(utts)
    for proc in (MfccProcessor(dither=0), FilterbankProcessor(dither=0),
                 MfccProcessor(dither=0, sample_rate=16000,
                               frame_length=0.05)):
        batch = proc.process_all(utterances, njobs=2)
        assert list(batch.keys()) == [u[0] for u in utts]
        for i, (name, _) in enumerate(utts):
            single = proc.process(Audio(synth_utterance(i, lengths[i]), 16000))
            assert batch[name] == single
            ref = oracle.features(
                proc.name, synth_utterance(i, lengths[i]),
                frame_length=float(proc.frame_length))
            scale_close(batch[name].data, ref, tol=1e-4)


def test_stability(pcm):
    """same processor twice / two fresh processors give == features at
    dither 0 (test/processor/test_stability.py:32-62)"""
    audio = Audio(pcm, 16000)
    for cls in (MfccProcessor, FilterbankProcessor, PlpProcessor,
                SpectrogramProcessor):
        p = cls(dither=0)
        assert p.process(audio) == p.process(audio) == cls(
            dither=0).process(audio)


def test_dither_is_distributional(pcm):
    """default dither=1.0: features are random but close to the dither-free
    ones, different at each call, and the noise has the right variance"""
    audio = Audio(pcm, 16000)
    clean = FilterbankProcessor(dither=0).process(audio).data
    a = FilterbankProcessor().process(audio).data
    b = FilterbankProcessor().process(audio).data
    assert not np.array_equal(a, b)
    assert np.abs(a - clean).max() < 1.5 and np.abs(a - clean).mean() < 0.05
    # pure noise: on a zero signal with rectangular window, no pre-emphasis,
    # no DC removal, E|X_k|^2 = W * dither^2 for every bin
    zeros = Audio(np.zeros(160000, np.int16), 16000)
    spec = SpectrogramProcessor(
        dither=2.0, window_type='rectangular', preemph_coeff=0,
        remove_dc_offset=False).process(zeros).data
    power = np.exp(spec[:, 1:256].astype(np.float64))
    assert abs(power.mean() / (400 * 4.0) - 1) < 0.02
    energy = EnergyProcessor(dither=2.0, compression='off',
                             remove_dc_offset=False).process(zeros).data
    assert abs(energy.mean() / (400 * 4.0) - 1) < 0.02


def test_full_size_properties():
    """BASELINE configs[1] sized batch slice (64 x 10 s): structural
    properties that do not need the oracle at full size"""
    from shennong_b200 import engine, _lib
    signals = [synth_utterance(i) for i in range(64)]
    proc = FilterbankProcessor(dither=0, num_bins=40)
    plan = engine.feature_plan(proc._frame_opts(), proc._mel_opts(),
                               proc._feat_opts())
    assert plan.fast_path
    packed = engine.PackedAudio(signals)
    batch = engine.Batch(plan, packed)
    assert batch.total_frames == 64 * 998
    out = engine.to_host(engine.compute_features(plan, batch))
    assert out.shape == (64 * 998, 40) and np.isfinite(out).all()
    # permutation equivariance: reversing the batch reverses the blocks
    packed2 = engine.PackedAudio(signals[::-1])
    out2 = engine.to_host(engine.compute_features(
        plan, engine.Batch(plan, packed2)))
    assert np.array_equal(out2.reshape(64, 998, 40)[::-1],
                          out.reshape(64, 998, 40))
    # spot-check three utterances against the oracle
    for i in (0, 31, 63):
        scale_close(out[i * 998:(i + 1) * 998],
                    oracle.features('filterbank', signals[i], num_bins=40))
    # linearity of the un-logged magnitude path: scaling PCM by 2 scales the
    # linear filterbank by 4
    lin = FilterbankProcessor(dither=0, use_log_fbank=False)
    half = (signals[0] // 2).astype(np.int16)
    a = lin.process(Audio(half * 2, 16000)).data
    b = lin.process(Audio(half, 16000)).data
    assert np.allclose(a, 4 * b, rtol=1e-5)


def test_element_gather_path_without_tma():
    """The cooperative (non bulk-copy) staging of the fused kernel -- used for
    unaligned PCM buffers and utterance edges -- on every tile: the golden,
    ragged-batch and edge-length tests again in a process with SNB_NO_TMA=1"""
    import os
    import subprocess
    import sys
    if os.environ.get('SNB_NO_TMA'):
        pytest.skip('already inside the SNB_NO_TMA run')
    env = dict(os.environ, SNB_NO_TMA='1')
    here = os.path.abspath(__file__)
    res = subprocess.run(
        [sys.executable, '-m', 'pytest', here, '-m', 'gpu', '-q', '-x', '-k',
         'golden or ragged or edge_lengths or plp_py'],
        env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_unaligned_utterance_offsets():
    """The C ABI takes any sample_begin: utterances packed at odd / non 16-byte
    offsets go through the misaligned bulk-copy and the element-wise load
    paths of the fused kernel and must give the same rows"""
    import torch
    from shennong_b200 import engine
    lengths = [22713, 5000, 16001, 48000, 401]
    sigs = [synth_utterance(10 + i, n) for i, n in enumerate(lengths)]
    gaps = [3, 1, 7, 2, 5]                    # odd and even misalignments
    starts, pos = [], 0
    for n, gap in zip(lengths, gaps):
        pos += gap
        starts.append(pos)
        pos += n
    buf = np.zeros(pos + 64, dtype=np.int16)
    for s, start in zip(sigs, starts):
        buf[start:start + len(s)] = s
    dev = torch.from_numpy(buf).cuda()
    for proc in (MfccProcessor(dither=0), FilterbankProcessor(dither=0, num_bins=40),
                 MfccProcessor(dither=0, snip_edges=False)):
        plan = engine.feature_plan(
            proc._frame_opts(), proc._mel_opts(), proc._feat_opts())
        packed = engine.PackedAudio.from_packed(
            None, np.array(starts), np.array(lengths), dev=dev)
        batch = engine.Batch(plan, packed)
        out = engine.to_host(engine.compute_features(plan, batch))
        offs = batch.frame_offsets
        for i, sig in enumerate(sigs):
            single = proc.process(Audio(sig, 16000)).data
            assert np.array_equal(out[offs[i]:offs[i + 1]], single), i


def test_dither_statistics():
    """dither = N(0, 1) per sample of every extracted window, drawn anew for
    each frame (Kaldi dithers the frame's private copy, overlapping frames do
    not share their noise): on a silent signal the frame energy is a
    chi-square with 400 degrees of freedom -- mean 400, variance 800 -- and
    the energies of different frames are uncorrelated, on the fused fast path
    (counter-based Box-Muller) as on the generic one"""
    from shennong_b200.processor import EnergyProcessor
    silence = Audio(np.zeros(16000 * 200, np.int16), 16000)
    for kwargs in ({}, {'frame_length': 0.04}):        # fast path, generic path
        proc = EnergyProcessor(
            dither=1.0, remove_dc_offset=False, raw_energy=True,
            compression='off', **kwargs)
        e = proc.process(silence).data.reshape(-1)
        dof = int(16000 * proc.frame_length)
        assert e.shape[0] > 19000
        assert abs(e.mean() / dof - 1.0) < 0.01
        assert abs(e.var() / (2 * dof) - 1.0) < 0.1
        for lag in (1, 2, 3, 7):
            c = np.corrcoef(e[:-lag], e[lag:])[0, 1]
            assert abs(c) < 0.05, (lag, c)
        # another seed gives other noise
        e2 = proc.process(silence).data.reshape(-1)
        assert abs(np.corrcoef(e, e2)[0, 1]) < 0.05
