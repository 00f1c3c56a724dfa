"""ctypes binding of libsnb.so (the C ABI declared in include/snb.h)

This is the only place the Python host talks to native code: it plays the
role pykaldi's CLIF wrappers play in the reference.  The library is built
in-tree by ``shennong_b200/csrc/build.sh`` (see ``__graft_entry__.build``).
There is NO CPU fallback: if the library is missing the import of any
processor fails loudly.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'libsnb.so')

SNB_OK = 0
SNB_ERR_OPTION = -1
SNB_ERR_VALUE = -2
SNB_ERR_CUDA = -3
SNB_ERR_UNSUPPORTED = -4

WINDOW_TYPES = {'hamming': 0, 'hanning': 1, 'povey': 2, 'rectangular': 3,
                'blackman': 4}
FEATURE_KINDS = {'spectrogram': 0, 'filterbank': 1, 'mfcc': 2, 'plp': 3,
                 'energy': 4}
ENERGY_COMPRESSION = {'off': 0, 'log': 1, 'sqrt': 2}

i32, i64, f32, u64 = (ctypes.c_int32, ctypes.c_int64, ctypes.c_float,
                      ctypes.c_uint64)
vp = ctypes.c_void_p


class FrameOpts(ctypes.Structure):
    """snb_frame_opts"""
    _fields_ = [('samp_freq', f32), ('frame_shift_ms', f32),
                ('frame_length_ms', f32), ('dither', f32),
                ('preemph_coeff', f32), ('blackman_coeff', f32),
                ('remove_dc_offset', i32), ('window_type', i32),
                ('round_to_power_of_two', i32), ('snip_edges', i32)]


class MelOpts(ctypes.Structure):
    """snb_mel_opts"""
    _fields_ = [('num_bins', i32), ('low_freq', f32), ('high_freq', f32),
                ('vtln_low', f32), ('vtln_high', f32)]


class FeatOpts(ctypes.Structure):
    """snb_feat_opts"""
    _fields_ = [('kind', i32), ('num_ceps', i32), ('use_energy', i32),
                ('energy_floor', f32), ('raw_energy', i32),
                ('cepstral_lifter', f32), ('htk_compat', i32),
                ('use_log_fbank', i32), ('use_power', i32),
                ('lpc_order', i32), ('compress_factor', f32),
                ('cepstral_scale', f32), ('rasta', i32),
                ('energy_compression', i32)]


class PitchOpts(ctypes.Structure):
    """snb_pitch_opts"""
    _fields_ = [('samp_freq', f32), ('frame_shift_ms', f32),
                ('frame_length_ms', f32), ('preemph_coeff', f32),
                ('min_f0', f32), ('max_f0', f32), ('soft_min_f0', f32),
                ('penalty_factor', f32), ('lowpass_cutoff', f32),
                ('resample_freq', f32), ('delta_pitch', f32),
                ('nccf_ballast', f32), ('lowpass_filter_width', i32),
                ('upsample_filter_width', i32), ('snip_edges', i32)]


class PitchPostOpts(ctypes.Structure):
    """snb_pitch_post_opts"""
    _fields_ = [('pitch_scale', f32), ('pov_scale', f32), ('pov_offset', f32),
                ('delta_pitch_scale', f32),
                ('delta_pitch_noise_stddev', f32),
                ('normalization_left_context', i32),
                ('normalization_right_context', i32),
                ('delta_window', i32), ('delay', i32),
                ('add_pov_feature', i32), ('add_normalized_log_pitch', i32),
                ('add_delta_pitch', i32), ('add_raw_log_pitch', i32)]


def struct_key(struct):
    """Hashable identity of a POD options struct (its raw bytes)"""
    return bytes(memoryview(struct))


# name -> (restype, argtypes): every symbol include/snb.h declares
SIGNATURES = {
    'snb_version': (ctypes.c_int, []),
    'snb_last_error': (ctypes.c_char_p, []),
    'snb_launch_count': (i64, []),
    'snb_window_size': (i32, [vp]),
    'snb_window_shift': (i32, [vp]),
    'snb_padded_window_size': (i32, [vp]),
    'snb_num_frames': (i64, [i64, vp]),
    'snb_first_sample_of_frame': (i64, [i32, vp]),
    'snb_window_function': (ctypes.c_int, [vp, vp, i32]),
    'snb_mel_banks_host': (ctypes.c_int, [vp, vp, f32, vp, vp]),
    'snb_feature_plan_create': (ctypes.c_int, [vp, vp, vp, vp]),
    'snb_pitch_plan_create': (ctypes.c_int, [vp, vp]),
    'snb_plan_destroy': (None, [vp]),
    'snb_plan_dim': (i32, [vp]),
    'snb_plan_uses_fast_path': (i32, [vp]),
    'snb_batch_create': (ctypes.c_int, [vp, vp, vp, i64, vp, vp]),
    'snb_batch_create_on_stream': (ctypes.c_int,
                                   [vp, vp, vp, i64, vp, vp, vp]),
    'snb_batch_destroy': (None, [vp]),
    'snb_batch_num_utts': (i64, [vp]),
    'snb_batch_total_frames': (i64, [vp]),
    'snb_batch_frame_offsets': (vp, [vp]),
    'snb_batch_frame_offsets_device': (vp, [vp]),
    'snb_compute_features': (ctypes.c_int, [vp, vp, vp, i64, u64, vp, i64,
                                            vp]),
    'snb_feature_workspace_bytes': (i64, [vp, vp]),
    'snb_compute_features_ws': (ctypes.c_int, [vp, vp, vp, i64, u64, vp, i64,
                                               vp, i64, vp]),
    'snb_compute_features_f32': (ctypes.c_int, [vp, vp, vp, i64, u64, vp, i64,
                                                vp]),
    'snb_compute_deltas': (ctypes.c_int, [vp, i64, i32, vp, i64, i64, i32,
                                          i32, vp, i64, vp]),
    'snb_cmvn_accumulate': (ctypes.c_int, [vp, i64, i32, vp, i64, vp, vp,
                                           vp]),
    'snb_cmvn_reduce_groups': (ctypes.c_int, [vp, i32, vp, vp, i64, vp, vp]),
    'snb_cmvn_norm_from_stats': (ctypes.c_int, [vp, i64, i32, i32, i32, vp,
                                                vp]),
    'snb_cmvn_apply': (ctypes.c_int, [vp, i64, i32, vp, i64, i64, vp, vp, vp,
                                      i64, vp]),
    'snb_cmvn_apply_deltas': (ctypes.c_int, [vp, i64, i32, vp, i64, i64, vp,
                                             vp, i32, i32, vp, i64, vp]),
    'snb_sliding_window_cmn': (ctypes.c_int, [vp, i64, i32, vp, i64, i64, i32,
                                              i32, i32, i32, vp, i64, vp]),
    'snb_vad_energy': (ctypes.c_int, [vp, i64, vp, i64, i64, f32, f32, i32,
                                      f32, vp, vp]),
    'snb_convert_f64_to_f32': (ctypes.c_int, [vp, vp, i64, vp]),
    'snb_pitch_num_frames': (i64, [i64, vp]),
    'snb_pitch_num_frames_array': (None, [vp, i64, vp, vp]),
    'snb_pitch_wave_utts': (i64, [vp]),
    'snb_pitch_workspace_bytes': (i64, [vp, vp]),
    'snb_compute_pitch': (ctypes.c_int, [vp, vp, vp, vp, i64, vp, i64, vp]),
    'snb_process_pitch_dim': (i32, [vp]),
    'snb_process_pitch': (ctypes.c_int, [vp, vp, i64, vp, vp, i64, i64, i64,
                                         u64, vp, i64, vp]),
    'snb_peer_buffer_create': (ctypes.c_int, [i64, vp, vp]),
    'snb_peer_buffer_open': (ctypes.c_int, [vp, vp]),
    'snb_peer_buffer_close': (ctypes.c_int, [vp]),
    'snb_peer_buffer_destroy': (ctypes.c_int, [vp]),
    'snb_gather_rows': (ctypes.c_int, [vp, i64, vp, i32, i64, i32, vp]),
    'snb_gather_rows_bulk': (ctypes.c_int, [vp, i64, vp, i32, i64, i32, vp]),
    'snb_gather_rows_ce': (ctypes.c_int, [vp, i64, vp, i32, i64, vp]),
    'snb_wav_scan_batch': (ctypes.c_int, [vp, i64, vp, vp, vp, i32]),
    'snb_resampler_create': (ctypes.c_int, [i32, i32, ctypes.c_float, i32, vp]),
    'snb_resampler_destroy': (None, [vp]),
    'snb_resampler_num_out': (i64, [vp, i64]),
    'snb_resample_batch': (ctypes.c_int, [vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp]),
    'snb_read_segments': (ctypes.c_int, [vp, vp, vp, vp, i64, i32, vp]),
}

_lib = None


def lib():
    """Loads libsnb.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} not found: build the CUDA extension first '
                f'(python -c "import __graft_entry__ as g; g.build()" or '
                f'bash shennong_b200/csrc/build.sh). There is no CPU '
                f'fallback for the feature extraction hot path.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fun = getattr(handle, name)
            fun.restype = restype
            fun.argtypes = argtypes
        _lib = handle
    return _lib


def last_error():
    return (lib().snb_last_error() or b'').decode()


def check(code):
    """Maps a status code to the exception type the reference raises"""
    if code == SNB_OK:
        return
    message = last_error()
    if code == SNB_ERR_VALUE:
        raise ValueError(message)
    if code == SNB_ERR_UNSUPPORTED:
        raise NotImplementedError(message)
    # SNB_ERR_OPTION mirrors Kaldi's KALDI_ERR, surfaced by pykaldi as
    # RuntimeError (test/processor/test_mfcc.py:69-97)
    raise RuntimeError(message)


def ref(struct):
    return ctypes.cast(ctypes.pointer(struct), vp)


def np_ptr(array):
    return array.ctypes.data_as(vp)


def make_frame_opts(sample_rate, frame_shift_ms, frame_length_ms, dither,
                    preemph_coeff, remove_dc_offset, window_type,
                    round_to_power_of_two, blackman_coeff, snip_edges):
    return FrameOpts(
        np.float32(sample_rate), np.float32(frame_shift_ms),
        np.float32(frame_length_ms), np.float32(dither),
        np.float32(preemph_coeff), np.float32(blackman_coeff),
        int(bool(remove_dc_offset)), WINDOW_TYPES[window_type],
        int(bool(round_to_power_of_two)), int(bool(snip_edges)))
