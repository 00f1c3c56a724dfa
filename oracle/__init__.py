"""CPU oracle for the shennong frame-based feature hot path (ctypes front-end).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and only as the checker.  The product (``shennong_b200``) never
imports it and fails loudly when its CUDA library is missing.

The arithmetic lives in ``kaldi_oracle.c`` (see ``kaldi_oracle.h`` for
provenance and the reference file:line each function follows).  Parity
pinning status is documented in ``oracle/README.md``.
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'libkaldi_oracle.so')
_lib = None

WINDOWS = {'hamming': 0, 'hanning': 1, 'povey': 2, 'rectangular': 3,
           'blackman': 4}
KINDS = {'spectrogram': 0, 'filterbank': 1, 'mfcc': 2, 'plp': 3, 'energy': 4}
COMPRESSION = {'off': 0, 'log': 1, 'sqrt': 2}


class FrameOpts(ctypes.Structure):
    _fields_ = [('samp_freq', ctypes.c_float),
                ('frame_shift_ms', ctypes.c_float),
                ('frame_length_ms', ctypes.c_float),
                ('dither', ctypes.c_float),
                ('preemph_coeff', ctypes.c_float),
                ('blackman_coeff', ctypes.c_float),
                ('remove_dc_offset', ctypes.c_int32),
                ('window_type', ctypes.c_int32),
                ('round_to_power_of_two', ctypes.c_int32),
                ('snip_edges', ctypes.c_int32)]


class MelOpts(ctypes.Structure):
    _fields_ = [('num_bins', ctypes.c_int32),
                ('low_freq', ctypes.c_float),
                ('high_freq', ctypes.c_float),
                ('vtln_low', ctypes.c_float),
                ('vtln_high', ctypes.c_float)]


class FeatOpts(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32),
                ('num_ceps', ctypes.c_int32),
                ('use_energy', ctypes.c_int32),
                ('energy_floor', ctypes.c_float),
                ('raw_energy', ctypes.c_int32),
                ('cepstral_lifter', ctypes.c_float),
                ('htk_compat', ctypes.c_int32),
                ('use_log_fbank', ctypes.c_int32),
                ('use_power', ctypes.c_int32),
                ('lpc_order', ctypes.c_int32),
                ('compress_factor', ctypes.c_float),
                ('cepstral_scale', ctypes.c_float),
                ('rasta', ctypes.c_int32),
                ('energy_compression', ctypes.c_int32)]


class PitchOpts(ctypes.Structure):
    _fields_ = [('samp_freq', ctypes.c_float),
                ('frame_shift_ms', ctypes.c_float),
                ('frame_length_ms', ctypes.c_float),
                ('preemph_coeff', ctypes.c_float),
                ('min_f0', ctypes.c_float),
                ('max_f0', ctypes.c_float),
                ('soft_min_f0', ctypes.c_float),
                ('penalty_factor', ctypes.c_float),
                ('lowpass_cutoff', ctypes.c_float),
                ('resample_freq', ctypes.c_float),
                ('delta_pitch', ctypes.c_float),
                ('nccf_ballast', ctypes.c_float),
                ('lowpass_filter_width', ctypes.c_int32),
                ('upsample_filter_width', ctypes.c_int32),
                ('snip_edges', ctypes.c_int32),
                ('recompute_frame', ctypes.c_int32)]


class PitchPostOpts(ctypes.Structure):
    _fields_ = [('pitch_scale', ctypes.c_float),
                ('pov_scale', ctypes.c_float),
                ('pov_offset', ctypes.c_float),
                ('delta_pitch_scale', ctypes.c_float),
                ('delta_pitch_noise_stddev', ctypes.c_float),
                ('normalization_left_context', ctypes.c_int32),
                ('normalization_right_context', ctypes.c_int32),
                ('delta_window', ctypes.c_int32),
                ('delay', ctypes.c_int32),
                ('add_pov_feature', ctypes.c_int32),
                ('add_normalized_log_pitch', ctypes.c_int32),
                ('add_delta_pitch', ctypes.c_int32),
                ('add_raw_log_pitch', ctypes.c_int32)]


def build(force=False):
    """Compiles the C oracle into oracle/_build/ (gcc, a few seconds)"""
    src = [os.path.join(_HERE, f) for f in ('kaldi_oracle.c', 'kaldi_oracle.h')]
    if (not force and os.path.isfile(_LIB_PATH) and
            all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s)
                for s in src)):
        return _LIB_PATH
    subprocess.run(['make', '-C', _HERE, '-B'], check=True,
                   stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        L = _lib
        i32, i64, f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
        P = ctypes.c_void_p
        for name in ('orc_window_size', 'orc_window_shift',
                     'orc_padded_window_size'):
            getattr(L, name).restype = i32
            getattr(L, name).argtypes = [P]
        L.orc_num_frames.restype = i64
        L.orc_num_frames.argtypes = [i64, P]
        L.orc_first_sample_of_frame.restype = i64
        L.orc_first_sample_of_frame.argtypes = [i32, P]
        L.orc_window_function.argtypes = [P, P]
        L.orc_mel_banks.restype = i32
        L.orc_mel_banks.argtypes = [P, P, f32, P, P]
        L.orc_feat_dim.restype = i32
        L.orc_feat_dim.argtypes = [P, P, P]
        L.orc_compute_features.restype = i64
        L.orc_compute_features.argtypes = [P, i64, P, P, P, f32, P, P]
        L.orc_compute_features_batch.restype = i64
        L.orc_compute_features_batch.argtypes = [P, P, P, i64, P, P, P, P, i32]
        L.orc_compute_deltas.argtypes = [P, i64, i32, i32, i32, P]
        L.orc_cmvn_accumulate.argtypes = [P, i64, i32, P, P]
        L.orc_cmvn_apply.restype = i32
        L.orc_cmvn_apply.argtypes = [P, i32, i32, i32, P, i64]
        L.orc_sliding_window_cmn.argtypes = [P, i64, i32, i32, i32, i32, i32, P]
        L.orc_resample_num_out.restype = i64
        L.orc_resample_num_out.argtypes = [i64, i32, i32]
        L.orc_resample.argtypes = [P, i64, i32, i32, f32, i32, P]
        L.orc_vad_energy.argtypes = [P, i64, i32, f32, f32, i32, f32, P]
        L.orc_pitch_num_frames.restype = i64
        L.orc_pitch_num_frames.argtypes = [i64, P]
        L.orc_pitch_num_lags.restype = i32
        L.orc_pitch_num_lags.argtypes = [P]
        L.orc_compute_kaldi_pitch.restype = i64
        L.orc_compute_kaldi_pitch.argtypes = [P, i64, P, P]
        L.orc_process_pitch_dim.restype = i32
        L.orc_process_pitch_dim.argtypes = [P]
        L.orc_process_pitch.restype = i64
        L.orc_process_pitch.argtypes = [P, i64, P, P]
        L.orc_pipeline_batch.restype = i64
        L.orc_pipeline_batch.argtypes = [
            P, P, P, i64, P, P, P, i32, i32, i32, i32, P, i32]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def frame_opts(sample_rate=16000, frame_shift=0.01, frame_length=0.025,
               dither=0.0, preemph_coeff=0.97, remove_dc_offset=True,
               window_type='povey', round_to_power_of_two=True,
               blackman_coeff=0.42, snip_edges=True, **_):
    # shennong setters store value * 1000.0 into Kaldi's float32 fields
    # (shennong/processor/base.py:162-172)
    return FrameOpts(
        np.float32(sample_rate), np.float32(frame_shift * 1000.0),
        np.float32(frame_length * 1000.0), np.float32(dither),
        np.float32(preemph_coeff), np.float32(blackman_coeff),
        int(bool(remove_dc_offset)), WINDOWS[window_type],
        int(bool(round_to_power_of_two)), int(bool(snip_edges)))


def mel_opts(num_bins=23, low_freq=20, high_freq=0, vtln_low=100,
             vtln_high=-500, **_):
    return MelOpts(int(num_bins), np.float32(low_freq), np.float32(high_freq),
                   np.float32(vtln_low), np.float32(vtln_high))


def feat_opts(kind, num_ceps=13, use_energy=None, energy_floor=0.0,
              raw_energy=True, cepstral_lifter=22.0, htk_compat=False,
              use_log_fbank=True, use_power=True, lpc_order=12,
              compress_factor=1.0 / 3.0, cepstral_scale=1.0, rasta=False,
              compression='log', **_):
    if use_energy is None:
        use_energy = kind in ('mfcc', 'plp')
    return FeatOpts(
        KINDS[kind], int(num_ceps), int(bool(use_energy)),
        np.float32(energy_floor), int(bool(raw_energy)),
        np.float32(cepstral_lifter), int(bool(htk_compat)),
        int(bool(use_log_fbank)), int(bool(use_power)), int(lpc_order),
        np.float32(compress_factor), np.float32(cepstral_scale),
        int(bool(rasta)), COMPRESSION[compression])


def pitch_opts(sample_rate=16000, frame_shift=0.01, frame_length=0.025,
               min_f0=50, max_f0=400, soft_min_f0=10, penalty_factor=0.1,
               lowpass_cutoff=1000, resample_freq=4000, delta_pitch=0.005,
               nccf_ballast=7000, lowpass_filter_width=1,
               upsample_filter_width=5, **_):
    return PitchOpts(
        np.float32(sample_rate), np.float32(frame_shift * 1000.0),
        np.float32(frame_length * 1000.0), 0.0, np.float32(min_f0),
        np.float32(max_f0), np.float32(soft_min_f0),
        np.float32(penalty_factor), np.float32(lowpass_cutoff),
        np.float32(resample_freq), np.float32(delta_pitch),
        np.float32(nccf_ballast), int(lowpass_filter_width),
        int(upsample_filter_width), 1, 500)


def pitch_post_opts(pitch_scale=2.0, pov_scale=2.0, pov_offset=0.0,
                    delta_pitch_scale=10.0, delta_pitch_noise_stddev=0.0,
                    normalization_left_context=75,
                    normalization_right_context=75, delta_window=2, delay=0,
                    add_pov_feature=True, add_normalized_log_pitch=True,
                    add_delta_pitch=True, add_raw_log_pitch=False, **_):
    return PitchPostOpts(
        np.float32(pitch_scale), np.float32(pov_scale),
        np.float32(pov_offset), np.float32(delta_pitch_scale),
        np.float32(delta_pitch_noise_stddev),
        int(normalization_left_context), int(normalization_right_context),
        int(delta_window), int(delay), int(bool(add_pov_feature)),
        int(bool(add_normalized_log_pitch)), int(bool(add_delta_pitch)),
        int(bool(add_raw_log_pitch)))


def to_int16_like_reference(data):
    """Audio.astype(np.int16) of the reference (shennong/audio.py:495-518)"""
    data = np.asarray(data)
    if data.dtype == np.int16:
        return data
    if data.dtype == np.int32:
        return (data / 2**15).astype(np.int16)
    return (data * 2**15).astype(np.int16)


def num_frames(nsamples, **kw):
    fo = frame_opts(**kw)
    return int(lib().orc_num_frames(int(nsamples), ctypes.byref(fo)))


def window(length, type='povey', blackman_coeff=0.42):
    fo = frame_opts(sample_rate=1000, frame_length=length / 1000.0,
                    window_type=type, blackman_coeff=blackman_coeff)
    out = np.zeros(lib().orc_window_size(ctypes.byref(fo)), dtype=np.float32)
    lib().orc_window_function(ctypes.byref(fo), _ptr(out))
    return out


def mel_banks(vtln_warp=1.0, **kw):
    fo, mo = frame_opts(**kw), mel_opts(**kw)
    nfft = lib().orc_padded_window_size(ctypes.byref(fo)) // 2
    w = np.zeros((max(mo.num_bins, 1), nfft), dtype=np.float32)
    c = np.zeros(max(mo.num_bins, 1), dtype=np.float32)
    ret = lib().orc_mel_banks(ctypes.byref(fo), ctypes.byref(mo),
                              np.float32(vtln_warp), _ptr(w), _ptr(c))
    if ret != 0:
        raise RuntimeError('invalid mel options')
    return w, c


def features(kind, signal, vtln_warp=1.0, **kw):
    """Oracle features of `kind` on a 1d signal.

    For all kinds but 'energy' the signal is cast to int16 exactly as the
    reference does (processor/base.py:428); energy keeps the raw scale
    (processor/energy.py:158).  Returns float32 [nframes, dim] (float64 for
    energy).
    """
    fo, mo, xo = frame_opts(**kw), mel_opts(**kw), feat_opts(kind, **kw)
    if kind == 'energy':
        wave = np.ascontiguousarray(np.asarray(signal), dtype=np.float32)
    else:
        wave = np.ascontiguousarray(
            to_int16_like_reference(signal), dtype=np.float32)
    L = lib()
    dim = L.orc_feat_dim(ctypes.byref(fo), ctypes.byref(mo), ctypes.byref(xo))
    nf = L.orc_num_frames(len(wave), ctypes.byref(fo))
    if dim <= 0 or nf < 0:
        raise RuntimeError('invalid options')
    out = np.zeros((nf, dim), dtype=np.float32)
    out64 = np.zeros((nf, 1), dtype=np.float64) if kind == 'energy' else None
    ret = L.orc_compute_features(
        _ptr(wave), len(wave), ctypes.byref(fo), ctypes.byref(mo),
        ctypes.byref(xo), np.float32(vtln_warp), _ptr(out), _ptr(out64))
    if ret < 0:
        raise RuntimeError('invalid options')
    return out64 if kind == 'energy' else out


def features_batch(kind, pcm, sample_offsets, nthreads=0, **kw):
    fo, mo, xo = frame_opts(**kw), mel_opts(**kw), feat_opts(kind, **kw)
    L = lib()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
    nf = np.array([L.orc_num_frames(int(so[i + 1] - so[i]), ctypes.byref(fo))
                   for i in range(len(so) - 1)], dtype=np.int64)
    fofs = np.concatenate(([0], np.cumsum(nf))).astype(np.int64)
    dim = L.orc_feat_dim(ctypes.byref(fo), ctypes.byref(mo), ctypes.byref(xo))
    out = np.zeros((int(fofs[-1]), dim), dtype=np.float32)
    L.orc_compute_features_batch(
        _ptr(pcm), _ptr(so), _ptr(fofs), len(so) - 1, ctypes.byref(fo),
        ctypes.byref(mo), ctypes.byref(xo), _ptr(out), int(nthreads))
    return out, fofs


def pipeline_batch(kind, pcm, sample_offsets, cmvn=True, norm_vars=True,
                   delta_order=2, delta_window=2, nthreads=0, **kw):
    fo, mo, xo = frame_opts(**kw), mel_opts(**kw), feat_opts(kind, **kw)
    L = lib()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
    nf = np.array([L.orc_num_frames(int(so[i + 1] - so[i]), ctypes.byref(fo))
                   for i in range(len(so) - 1)], dtype=np.int64)
    fofs = np.concatenate(([0], np.cumsum(nf))).astype(np.int64)
    dim = L.orc_feat_dim(ctypes.byref(fo), ctypes.byref(mo), ctypes.byref(xo))
    out = np.zeros((int(fofs[-1]), dim * (delta_order + 1)), dtype=np.float32)
    L.orc_pipeline_batch(
        _ptr(pcm), _ptr(so), _ptr(fofs), len(so) - 1, ctypes.byref(fo),
        ctypes.byref(mo), ctypes.byref(xo), int(cmvn), int(norm_vars),
        int(delta_order), int(delta_window), _ptr(out), int(nthreads))
    return out, fofs


def deltas(data, order=2, window=2):
    data = np.ascontiguousarray(data, dtype=np.float32)
    out = np.zeros((data.shape[0], data.shape[1] * (order + 1)), np.float32)
    lib().orc_compute_deltas(_ptr(data), data.shape[0], data.shape[1],
                             int(order), int(window), _ptr(out))
    return out


def cmvn_accumulate(data, weights=None, stats=None):
    data = np.ascontiguousarray(data, dtype=np.float32)
    dim = data.shape[1]
    if stats is None:
        stats = np.zeros((2, dim + 1), dtype=np.float64)
    stats = np.ascontiguousarray(stats, dtype=np.float64).copy()
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.float32)
    lib().orc_cmvn_accumulate(_ptr(data), data.shape[0], dim, _ptr(weights),
                              _ptr(stats))
    return stats


def cmvn_apply(data, stats, norm_vars=True, reverse=False, skip_dims=None):
    data = np.ascontiguousarray(data, dtype=np.float32).copy()
    stats = np.ascontiguousarray(stats, dtype=np.float64).copy()
    dim = data.shape[1]
    if skip_dims:
        # FakeStatsForSomeDims: mean 0, variance 1
        for d in skip_dims:
            stats[0, d] = 0.0
            stats[1, d] = stats[0, dim]
    ret = lib().orc_cmvn_apply(_ptr(stats), dim, int(norm_vars), int(reverse),
                               _ptr(data), data.shape[0])
    if ret != 0:
        raise ValueError('insufficient stats')
    return data


def sliding_window_cmn(data, center=True, cmn_window=600, min_window=100,
                       normalize_variance=False):
    data = np.ascontiguousarray(data, dtype=np.float32)
    out = np.zeros_like(data)
    lib().orc_sliding_window_cmn(
        _ptr(data), data.shape[0], data.shape[1], int(center),
        int(cmn_window), int(min_window), int(normalize_variance), _ptr(out))
    return out


def vad(data, energy_threshold=5.0, energy_mean_scale=0.5, frames_context=0,
        proportion_threshold=0.6):
    data = np.ascontiguousarray(data, dtype=np.float32)
    out = np.zeros(data.shape[0], dtype=np.float32)
    lib().orc_vad_energy(
        _ptr(data), data.shape[0], data.shape[1],
        np.float32(energy_threshold), np.float32(energy_mean_scale),
        int(frames_context), np.float32(proportion_threshold), _ptr(out))
    return out.astype(np.uint8).reshape(-1, 1)


def pitch(signal, **kw):
    po = pitch_opts(**kw)
    wave = np.ascontiguousarray(
        to_int16_like_reference(signal), dtype=np.float32)
    L = lib()
    nf = L.orc_pitch_num_frames(len(wave), ctypes.byref(po))
    out = np.zeros((nf, 2), dtype=np.float32)
    got = L.orc_compute_kaldi_pitch(_ptr(wave), len(wave), ctypes.byref(po),
                                    _ptr(out))
    assert got == nf, (got, nf)
    return out


def resample(signal, rate_in, rate_out, cutoff=0.0, num_zeros=0):
    """Kaldi's flushed LinearResample of a whole signal (float32 result)"""
    wave = np.ascontiguousarray(signal, dtype=np.float32)
    L = lib()
    n = L.orc_resample_num_out(len(wave), int(rate_in), int(rate_out))
    out = np.zeros(n, dtype=np.float32)
    L.orc_resample(_ptr(wave), len(wave), int(rate_in), int(rate_out),
                   float(cutoff), int(num_zeros), _ptr(out))
    return out


def process_pitch(raw, **kw):
    po = pitch_post_opts(**kw)
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    L = lib()
    dim = L.orc_process_pitch_dim(ctypes.byref(po))
    out = np.zeros((raw.shape[0] + po.delay, dim), dtype=np.float32)
    L.orc_process_pitch(_ptr(raw), raw.shape[0], ctypes.byref(po), _ptr(out))
    return out
