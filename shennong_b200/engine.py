"""Device-side engine: plans, ragged batches and kernel launches

PyTorch is used here ONLY as the device-memory container (tensors own the PCM,
feature and statistics buffers), for pinned host staging and for the current
CUDA stream.  Every numeric operation of the hot path is a kernel of
``libsnb.so`` reached through :mod:`shennong_b200._lib`.  There is no CPU
fallback: without a CUDA device every compute entry point raises.
"""

import ctypes
import threading

import numpy as np

from shennong_b200 import _lib

_ALIGN = 8          # utterances start on 16-byte boundaries in the packed PCM
_PAD = 64           # readable slack after the last sample (TMA rounds to 16 B)


def _torch():
    import torch
    return torch


def require_cuda():
    """Raises RuntimeError unless a CUDA device is usable"""
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError(
            'shennong_b200 needs a CUDA device (B200, sm_100a): the feature '
            'extraction hot path has no CPU fallback')
    return torch


def _stream_ptr():
    return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else None


# --------------------------------------------------------------------------
# plans (immutable, cached per device and option bytes)
# --------------------------------------------------------------------------
class _Plan:
    def __init__(self, handle, device):
        self.handle = handle
        self.device = device
        self.dim = int(_lib.lib().snb_plan_dim(handle))
        self.fast_path = bool(_lib.lib().snb_plan_uses_fast_path(handle))

    def __del__(self):
        try:
            _lib.lib().snb_plan_destroy(self.handle)
        except Exception:  # pragma: nocover (interpreter shutdown)
            pass


_plan_cache = {}
_plan_lock = threading.Lock()


def feature_plan(frame_opts, mel_opts, feat_opts):
    """Returns the (cached) plan for these option structs on this device"""
    torch = require_cuda()
    device = torch.cuda.current_device()
    key = ('feat', device, _lib.struct_key(frame_opts),
           _lib.struct_key(mel_opts) if mel_opts is not None else b'',
           _lib.struct_key(feat_opts))
    with _plan_lock:
        plan = _plan_cache.get(key)
        if plan is None:
            handle = ctypes.c_void_p()
            _lib.check(_lib.lib().snb_feature_plan_create(
                _lib.ref(frame_opts),
                _lib.ref(mel_opts) if mel_opts is not None else None,
                _lib.ref(feat_opts), ctypes.byref(handle)))
            plan = _Plan(handle, device)
            _plan_cache[key] = plan
    return plan


def pitch_plan(pitch_opts):
    torch = require_cuda()
    device = torch.cuda.current_device()
    key = ('pitch', device, _lib.struct_key(pitch_opts))
    with _plan_lock:
        plan = _plan_cache.get(key)
        if plan is None:
            handle = ctypes.c_void_p()
            _lib.check(_lib.lib().snb_pitch_plan_create(
                _lib.ref(pitch_opts), ctypes.byref(handle)))
            plan = _Plan(handle, device)
            _plan_cache[key] = plan
    return plan


def clear_plan_cache():
    with _plan_lock:
        _plan_cache.clear()


# --------------------------------------------------------------------------
# packed audio + ragged batch descriptors
# --------------------------------------------------------------------------
class PackedAudio:
    """PCM of several utterances packed in one pinned host buffer and its
    device copy; every utterance starts on a 16-byte boundary (TMA staging).

    dtype is int16 (what every processor of the reference feeds Kaldi,
    processor/base.py:428) or float32 (EnergyProcessor on float audio,
    energy.py:158).
    """

    def __init__(self, signals=None, dtype=np.int16):
        self.host = self.dev = None
        self.starts = self.lengths = None
        if signals is None:
            return
        torch = require_cuda()
        lengths = np.array([len(s) for s in signals], dtype=np.int64)
        padded = (lengths + _ALIGN - 1) // _ALIGN * _ALIGN
        starts = np.concatenate(([0], np.cumsum(padded)))[:-1].astype(np.int64)
        total = int(padded.sum()) + _PAD
        tdtype = torch.int16 if dtype == np.int16 else torch.float32
        self.host = torch.zeros(total, dtype=tdtype, pin_memory=True)
        view = self.host.numpy()
        for s, start in zip(signals, starts):
            view[start:start + len(s)] = s
        self.lengths = lengths
        self.starts = starts
        self.dev = self.host.to('cuda', non_blocking=True)

    @classmethod
    def from_packed(cls, host_tensor, starts, lengths, dev=None):
        """Wraps an already packed pinned tensor (bench end-to-end path);
        ``starts`` should be multiples of 8 samples."""
        self = cls()
        self.host = host_tensor
        self.starts = np.ascontiguousarray(starts, dtype=np.int64)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        self.dev = dev if dev is not None else host_tensor.to(
            'cuda', non_blocking=True)
        return self

    @property
    def nutts(self):
        return len(self.lengths)

    @property
    def is_float(self):
        return self.dev.dtype != _torch().int16


class Batch:
    """snb_batch handle: frame counts, device offsets, tile table"""

    def __init__(self, plan, packed, vtln_warps=None):
        L = _lib.lib()
        self.plan = plan
        self.packed = packed
        nutts = packed.nutts
        warps = None
        if vtln_warps is not None:
            warps = np.ascontiguousarray(vtln_warps, dtype=np.float32)
            if warps.shape != (nutts,):
                raise ValueError('one vtln warp per utterance is expected')
        handle = ctypes.c_void_p()
        # stream-ordered creation: the descriptor upload is queued on the
        # current stream (where the kernels that read it are launched) and
        # the device blob is recycled through the library's pool
        _lib.check(L.snb_batch_create_on_stream(
            plan.handle, _lib.np_ptr(packed.starts),
            _lib.np_ptr(packed.lengths), nutts,
            _lib.np_ptr(warps) if warps is not None else None,
            _stream_ptr(), ctypes.byref(handle)))
        self.handle = handle
        self.nutts = nutts
        self.total_frames = int(L.snb_batch_total_frames(handle))
        ptr = L.snb_batch_frame_offsets(handle)
        self.frame_offsets = np.ctypeslib.as_array(
            ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)),
            shape=(nutts + 1,)).copy()
        self.d_frame_offsets = ctypes.c_void_p(
            L.snb_batch_frame_offsets_device(handle))

    @property
    def max_frames(self):
        if self.nutts == 0:
            return 0
        return int(np.diff(self.frame_offsets).max())

    def __del__(self):
        try:
            _lib.lib().snb_batch_destroy(self.handle)
        except Exception:  # pragma: nocover
            pass


# --------------------------------------------------------------------------
# kernel launches (all asynchronous on torch's current stream)
# --------------------------------------------------------------------------
# optional (start, end) CUDA events recorded around the NEXT feature launch
# on torch's current stream (bench.py times the dominant kernel live, inside
# the timed step, with them)
feature_events = None


def compute_features(plan, batch, seed=0, out=None, float64=False):
    """[total_frames, dim] features of the batch (device tensor)"""
    global feature_events
    torch = require_cuda()
    events, feature_events = feature_events, None
    if events is not None:
        events[0].record()
        try:
            return compute_features(plan, batch, seed, out, float64)
        finally:
            events[1].record()
    dtype = torch.float64 if float64 else torch.float32
    if out is None:
        out = torch.empty((batch.total_frames, plan.dim), dtype=dtype,
                          device='cuda')
    packed = batch.packed
    L = _lib.lib()
    seed = ctypes.c_uint64(int(seed) & (2**64 - 1))
    if packed.is_float:
        _lib.check(L.snb_compute_features_f32(
            plan.handle, batch.handle, _ptr(packed.dev), packed.dev.numel(),
            seed, _ptr(out), out.stride(0), _stream_ptr()))
        return out
    nbytes = int(L.snb_feature_workspace_bytes(plan.handle, batch.handle))
    work = (torch.empty(nbytes, dtype=torch.uint8, device='cuda')
            if nbytes > 0 else None)      # RASTA-PLP scratch
    _lib.check(L.snb_compute_features_ws(
        plan.handle, batch.handle, _ptr(packed.dev), packed.dev.numel(),
        seed, _ptr(out), out.stride(0), _ptr(work), nbytes, _stream_ptr()))
    return out


def compute_pitch(plan, batch, out=None):
    torch = require_cuda()
    L = _lib.lib()
    if out is None:
        out = torch.empty((batch.total_frames, 2), dtype=torch.float32,
                          device='cuda')
    nbytes = int(L.snb_pitch_workspace_bytes(plan.handle, batch.handle))
    work = torch.empty(max(nbytes, 16), dtype=torch.uint8, device='cuda')
    _lib.check(L.snb_compute_pitch(
        plan.handle, batch.handle, _ptr(batch.packed.dev), _ptr(work),
        nbytes, _ptr(out), out.stride(0), _stream_ptr()))
    return out


class RowLayout:
    """Frame offsets of a packed [total_frames, d] matrix on the device"""

    def __init__(self, frame_offsets=None, batch=None):
        torch = require_cuda()
        self._batch = batch     # keeps the device array alive
        if batch is not None:
            self.ptr = batch.d_frame_offsets
            self.nutts = batch.nutts
            self.total = batch.total_frames
            self.max_frames = batch.max_frames
        else:
            offs = np.ascontiguousarray(frame_offsets, dtype=np.int64)
            self._dev = torch.from_numpy(offs).to('cuda')
            self.ptr = _ptr(self._dev)
            self.nutts = len(offs) - 1
            self.total = int(offs[-1])
            self.max_frames = int(np.diff(offs).max()) if len(offs) > 1 else 0


def deltas(x, layout, order, window, norm=None, utt_group=None, out=None):
    """Deltas (optionally fused with a CMVN apply) of x [total, d]"""
    torch = require_cuda()
    dim = x.shape[1]
    if out is None:
        out = torch.empty((x.shape[0], dim * (order + 1)),
                          dtype=torch.float32, device='cuda')
    _lib.check(_lib.lib().snb_cmvn_apply_deltas(
        _ptr(x), x.stride(0), dim, layout.ptr, layout.nutts, layout.total,
        _ptr(norm), _ptr(utt_group), int(order), int(window), _ptr(out),
        out.stride(0), _stream_ptr()))
    return out


def cmvn_accumulate(x, layout, weights=None, out=None):
    """Per-utterance CMVN statistics, float64 [nutts, 2, dim+1]"""
    torch = require_cuda()
    dim = x.shape[1]
    stats = out
    if stats is None:
        stats = torch.empty((layout.nutts, 2, dim + 1), dtype=torch.float64,
                            device='cuda')
    elif (tuple(stats.shape) != (layout.nutts, 2, dim + 1)
          or not stats.is_contiguous()):
        raise ValueError('bad statistics buffer')
    _lib.check(_lib.lib().snb_cmvn_accumulate(
        _ptr(x), x.stride(0), dim, layout.ptr, layout.nutts, _ptr(weights),
        _ptr(stats), _stream_ptr()))
    return stats


def cmvn_reduce_groups(utt_stats, group_ptr, group_utts, ngroups):
    """Deterministic per-group sums of per-utterance statistics"""
    torch = require_cuda()
    dim = utt_stats.shape[2] - 1
    out = torch.zeros((ngroups, 2, dim + 1), dtype=torch.float64,
                      device='cuda')
    gp = torch.from_numpy(np.ascontiguousarray(group_ptr, np.int64)).cuda()
    gu = torch.from_numpy(np.ascontiguousarray(group_utts, np.int64)).cuda()
    _lib.check(_lib.lib().snb_cmvn_reduce_groups(
        _ptr(utt_stats), dim, _ptr(gp), _ptr(gu), ngroups, _ptr(out),
        _stream_ptr()))
    return out


def cmvn_norm(stats, norm_vars=True, reverse=False, out=None):
    """float32 [ngroups, 2, dim] (offset, scale) table from float64 stats"""
    torch = require_cuda()
    ngroups, dim = stats.shape[0], stats.shape[2] - 1
    norm = out if out is not None else torch.empty(
        (ngroups, 2, dim), dtype=torch.float32, device='cuda')
    _lib.check(_lib.lib().snb_cmvn_norm_from_stats(
        _ptr(stats), ngroups, dim, int(bool(norm_vars)), int(bool(reverse)),
        _ptr(norm), _stream_ptr()))
    return norm


def cmvn_apply(x, layout, norm, utt_group=None, out=None):
    torch = require_cuda()
    dim = x.shape[1]
    if out is None:
        out = torch.empty((x.shape[0], dim), dtype=torch.float32,
                          device='cuda')
    _lib.check(_lib.lib().snb_cmvn_apply(
        _ptr(x), x.stride(0), dim, layout.ptr, layout.nutts, layout.total,
        _ptr(norm), _ptr(utt_group), _ptr(out), out.stride(0),
        _stream_ptr()))
    return out


def sliding_window_cmn(x, layout, center, cmn_window, min_window,
                       normalize_variance):
    torch = require_cuda()
    out = torch.empty_like(x)
    _lib.check(_lib.lib().snb_sliding_window_cmn(
        _ptr(x), x.stride(0), x.shape[1], layout.ptr, layout.nutts,
        layout.total, int(bool(center)), int(cmn_window), int(min_window),
        int(bool(normalize_variance)), _ptr(out), out.stride(0),
        _stream_ptr()))
    return out


def vad_energy(x, layout, energy_threshold, energy_mean_scale, frames_context,
               proportion_threshold):
    """0/1 float32 [total] voicing decisions from column 0 of x"""
    torch = require_cuda()
    out = torch.zeros(x.shape[0], dtype=torch.float32, device='cuda')
    _lib.check(_lib.lib().snb_vad_energy(
        _ptr(x), x.stride(0), layout.ptr, layout.nutts, layout.total,
        np.float32(energy_threshold), np.float32(energy_mean_scale),
        int(frames_context), np.float32(proportion_threshold), _ptr(out),
        _stream_ptr()))
    return out


def f64_to_f32(x):
    torch = require_cuda()
    out = torch.empty(x.shape, dtype=torch.float32, device='cuda')
    _lib.check(_lib.lib().snb_convert_f64_to_f32(
        _ptr(x), _ptr(out), x.numel(), _stream_ptr()))
    return out


def process_pitch(post_opts, raw, layout, seed=0, out=None, out_layout=None):
    """Pitch post-processing of raw [total, 2] (NCCF, pitch) rows

    With ``delay = d > 0`` an utterance of F > 0 frames gives F + d rows (row
    t holds the features of frame max(0, t - d), like Kaldi's ProcessPitch):
    the rows of utterance u then start at ``frame_offsets[u] + u * d`` and the
    output has ``total + nutts * d`` rows.  With `out_layout` (a RowLayout)
    the rows of utterance u go to its rows of `out` instead, trimmed to their
    number (pasting next to features that are a frame or two shorter).
    """
    torch = require_cuda()
    L = _lib.lib()
    dim = int(L.snb_process_pitch_dim(_lib.ref(post_opts)))
    delay = max(int(post_opts.delay), 0)
    if out is None:
        rows = (out_layout.total if out_layout is not None
                else raw.shape[0] + delay * layout.nutts)
        out = torch.zeros((rows, max(dim, 1)), dtype=torch.float32,
                          device='cuda')
    if out.shape[0] == 0 or raw.shape[0] == 0:
        return out          # (pitch frames only where the features have none)
    _lib.check(L.snb_process_pitch(
        _lib.ref(post_opts), _ptr(raw), raw.stride(0), layout.ptr,
        out_layout.ptr if out_layout is not None else None,
        layout.nutts, layout.total, layout.max_frames,
        ctypes.c_uint64(int(seed) & (2**64 - 1)), _ptr(out), out.stride(0),
        _stream_ptr()))
    return out


# --------------------------------------------------------------------------
# helpers for the host API
# --------------------------------------------------------------------------
class Resampler:
    """snb_resampler handle: Kaldi's LinearResample between two sample rates
    (cached per (rate_in, rate_out, cutoff, zeros) and device)"""
    _cache = {}
    _lock = threading.Lock()

    def __init__(self, rate_in, rate_out, lowpass_cutoff=0.0, num_zeros=0):
        require_cuda()
        self.rate_in, self.rate_out = int(rate_in), int(rate_out)
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().snb_resampler_create(
            self.rate_in, self.rate_out, float(lowpass_cutoff),
            int(num_zeros), ctypes.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.lib().snb_resampler_destroy(self.handle)
        except Exception:
            pass

    @classmethod
    def get(cls, rate_in, rate_out, lowpass_cutoff=0.0, num_zeros=0):
        torch = require_cuda()
        key = (int(rate_in), int(rate_out), float(lowpass_cutoff),
               int(num_zeros), torch.cuda.current_device())
        with cls._lock:
            if key not in cls._cache:
                cls._cache[key] = cls(rate_in, rate_out, lowpass_cutoff,
                                      num_zeros)
            return cls._cache[key]

    def num_out(self, nsamples):
        return int(_lib.lib().snb_resampler_num_out(self.handle, int(nsamples)))


def resample_packed(packed, rate_in, rate_out, lowpass_cutoff=0.0,
                    num_zeros=0, float32=False):
    """Sample-rate conversion of every utterance of an int16
    :class:`PackedAudio` on the device (``snb_resample_batch``)

    Returns a new PackedAudio at `rate_out` whose device buffer holds the
    int16 result (truncated toward zero like the reference's
    ``.astype(np.int16)``, audio.py:423), every utterance on a 16-byte
    boundary -- ready for the feature kernels, nothing goes through the host
    -- and, with `float32`, also the float32 device tensor of the same
    geometry (bit-equal to the CPU oracle).
    """
    torch = require_cuda()
    if packed.is_float:
        raise ValueError('int16 audio expected')
    rs = Resampler.get(rate_in, rate_out, lowpass_cutoff, num_zeros)
    lengths = np.array([rs.num_out(n) for n in packed.lengths], dtype=np.int64)
    padded = (lengths + _ALIGN - 1) // _ALIGN * _ALIGN
    starts = np.concatenate(([0], np.cumsum(padded)))[:-1].astype(np.int64)
    total = int(padded.sum()) + _PAD
    out = torch.zeros(total, dtype=torch.int16, device='cuda')
    outf = (torch.zeros(total, dtype=torch.float32, device='cuda')
            if float32 else None)
    nutts = packed.nutts
    step = 65535
    for b in range(0, nutts, step):
        e = min(b + step, nutts)
        desc = torch.from_numpy(np.concatenate(
            [packed.starts[b:e], packed.lengths[b:e], starts[b:e]])).to('cuda')
        counts = torch.empty(e - b, dtype=torch.int64, device='cuda')
        n = e - b
        _lib.check(_lib.lib().snb_resample_batch(
            rs.handle, _ptr(packed.dev), _ptr(desc[:n]), _ptr(desc[n:2 * n]),
            _ptr(desc[2 * n:]), n, int(lengths[b:e].max()) if n else 0,
            _ptr(outf), _ptr(out), _ptr(counts), _stream_ptr()))
    result = PackedAudio.from_packed(None, starts, lengths, dev=out)
    return (result, outf) if float32 else result


def num_frames_array(frame_opts, lengths):
    """Vectorised snb_num_frames (NumFrames with flush, frames.py:137) for an
    int64 array of utterance lengths"""
    L = _lib.lib()
    lengths = np.asarray(lengths, dtype=np.int64)
    size = int(L.snb_window_size(_lib.ref(frame_opts)))
    shift = int(L.snb_window_shift(_lib.ref(frame_opts)))
    if shift <= 0 or size <= 0:
        raise ValueError('cannot compute nframes: sample rate too low')
    if frame_opts.snip_edges:
        return np.where(lengths < size, 0, 1 + (lengths - size) // shift)
    return (lengths + shift // 2) // shift



def pitch_num_frames_array(pitch_opts, lengths):
    """snb_pitch_num_frames (frames compute_kaldi_pitch returns,
    pitch_kaldi.py:298) for an int64 array of utterance lengths"""
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    out = np.empty(len(lengths), dtype=np.int64)
    _lib.lib().snb_pitch_num_frames_array(
        _lib.np_ptr(lengths), len(lengths), _lib.ref(pitch_opts),
        _lib.np_ptr(out))
    return out


_seed_lock = threading.Lock()
_seed_state = np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0]


def next_seed():
    """A fresh 64-bit seed per call (dither is non reproducible in the
    reference too: Kaldi draws from libc rand())"""
    global _seed_state
    with _seed_lock:
        _seed_state = np.uint64(
            (int(_seed_state) * 6364136223846793005 + 1442695040888963407)
            % 2**64)
        return int(_seed_state)


def to_host(tensor):
    """Synchronous device -> host copy as a numpy array"""
    return tensor.detach().cpu().numpy()


def from_host(array, dtype=np.float32):
    torch = require_cuda()
    return torch.from_numpy(
        np.ascontiguousarray(array, dtype=dtype)).to('cuda')
