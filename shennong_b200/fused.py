"""Device-resident fused pipeline: features -> (VAD) -> CMVN -> delta ‖ pitch

This is the batched form of the reference's two-pass schedule
(shennong/pipeline.py:541-567, 570-648): everything between the packed int16
PCM and the final ``[frames, D]`` matrix stays in HBM and runs as a short
sequence of ``libsnb`` launches on one stream:

    snb_compute_features        base features            [F, d]
    snb_compute_features (energy) + snb_vad_energy        (with_vad only)
    snb_cmvn_accumulate         per-utterance f64 stats   [U, 2, d+1]
    snb_cmvn_reduce_groups      per-speaker stats         (by_speaker only)
    snb_cmvn_norm_from_stats    f32 (offset, scale)
    snb_cmvn_apply_deltas       normalise + deltas, written with the final
                                leading dimension (column block 0)
    snb_compute_pitch + snb_process_pitch   written into the last columns

so the concatenation of the reference (features.py:350-437) is a strided
write, not a copy.  ``run_host`` pipelines chunks of a pinned host buffer over
three streams (H2D / compute / D2H) for the end-to-end path.
"""

import numpy as np

from shennong_b200 import engine


class FusedPipeline:
    """Batched extraction of one pipeline configuration

    Parameters
    ----------
    processor : FramesProcessor
        The main features processor (mfcc, filterbank, plp, spectrogram)
    delta : DeltaPostProcessor or None
    cmvn : None, 'utterance' or 'speaker'
    norm_vars : bool
    vad : VadPostProcessor or None
        When given, CMVN statistics are weighted by the energy VAD
        (pipeline.py:588-596)
    energy : EnergyProcessor or None (required with `vad`)
    pitch : (KaldiPitchProcessor, KaldiPitchPostProcessor) or None
    """
    def __init__(self, processor, delta=None, cmvn=None, norm_vars=True,
                 vad=None, energy=None, pitch=None):
        self.processor = processor
        self.delta = delta
        self.cmvn = cmvn
        self.norm_vars = norm_vars
        self.vad = vad
        self.energy = energy
        self.pitch = pitch
        if vad is not None and energy is None:
            raise ValueError('vad needs an energy processor')
        if pitch is not None and pitch[1].delay != 0:
            # Kaldi returns nframes + delay rows: the reference fails on the
            # mismatch with the frame times (pitch_kaldi.py:535-540)
            raise ValueError(
                'pitch postprocessing with delay != 0 changes the number of '
                'frames: mismatch between data and times')
        self.base_dim = processor.ndims
        order = delta.order if delta is not None else 0
        self.feat_dim = self.base_dim * (order + 1)
        self.pitch_dim = pitch[1].ndims if pitch is not None else 0
        self.out_dim = self.feat_dim + self.pitch_dim

    # -- plans ---------------------------------------------------------------
    def _plans(self):
        p = self.processor
        plans = {'feat': engine.feature_plan(
            p._frame_opts(), p._mel_opts(), p._feat_opts())}
        if self.vad is not None:
            e = self.energy
            plans['energy'] = engine.feature_plan(
                e._frame_opts(), None, e._feat_opts())
        if self.pitch is not None:
            plans['pitch'] = engine.pitch_plan(self.pitch[0]._pitch_opts())
        return plans

    def run_device(self, packed, speakers=None, warps=None, seed=None,
                   out=None, plans=None, base_buf=None, batch=None):
        """Runs the whole pipeline on a PackedAudio already on the device

        Returns (out [total_frames, out_dim] device tensor, frame_offsets
        int64 [U+1] host array, stats float64 device tensor or None,
        group index of each utterance or None).  All launches are queued on
        the current stream; nothing synchronises.  `base_buf` is an optional
        preallocated [>= total_frames, base_dim] float32 buffer for the
        intermediate base features (chunked pipelines reuse it); `batch` an
        engine.Batch of the features plan already created for `packed`.
        """
        torch = engine.require_cuda()
        plans = plans or self._plans()
        p = self.processor
        if batch is None:
            batch = engine.Batch(plans['feat'], packed, warps)
        layout = engine.RowLayout(batch=batch)
        if seed is None:
            seed = engine.next_seed() if p.dither != 0 else 0
        total = batch.total_frames
        if out is None:
            out = torch.empty((total, self.out_dim), dtype=torch.float32,
                              device='cuda')
        simple = (self.cmvn is None and self.delta is None)
        if simple:
            base = out[:, :self.base_dim]
            engine.compute_features(plans['feat'], batch, seed=seed, out=base)
        else:
            base = engine.compute_features(
                plans['feat'], batch, seed=seed,
                out=None if base_buf is None else base_buf[:total])
        stats, utt_group, norm = None, None, None
        if self.cmvn is not None:
            weights = None
            if self.vad is not None:
                ebatch = engine.Batch(plans['energy'], packed)
                eseed = engine.next_seed() if self.energy.dither != 0 else 0
                e64 = engine.compute_features(
                    plans['energy'], ebatch, seed=eseed, float64=True)
                e32 = engine.f64_to_f32(e64)
                v = self.vad
                weights = engine.vad_energy(
                    e32, engine.RowLayout(batch=ebatch), v.energy_threshold,
                    v.energy_mean_scale, v.frames_context,
                    v.proportion_threshold)
                self._keep = (ebatch,)
            stats = engine.cmvn_accumulate(base, layout, weights)
            if self.cmvn == 'speaker':
                if speakers is None:
                    raise ValueError('speakers are required for cmvn by '
                                     'speaker')
                names = sorted(set(speakers))
                index = {s: i for i, s in enumerate(names)}
                group = np.array([index[s] for s in speakers], dtype=np.int64)
                order = np.argsort(group, kind='stable')
                ptr = np.concatenate(
                    ([0], np.cumsum(np.bincount(group, minlength=len(names)))))
                stats = engine.cmvn_reduce_groups(
                    stats, ptr, order, len(names))
                utt_group = torch.from_numpy(
                    group.astype(np.int32)).to('cuda')
                self._group_names = names
            norm = engine.cmvn_norm(stats, self.norm_vars, False)
        if not simple:
            order = self.delta.order if self.delta is not None else 0
            window = self.delta.window if self.delta is not None else 1
            engine.deltas(base, layout, order, window, norm=norm,
                          utt_group=utt_group, out=out[:, :self.feat_dim])
        if self.pitch is not None:
            pproc, ppost = self.pitch
            pbatch = engine.Batch(plans['pitch'], packed)
            if not np.array_equal(pbatch.frame_offsets, batch.frame_offsets):
                raise NotImplementedError(
                    'features and pitch have a different number of frames: '
                    'use the per-utterance API (Features.concatenate trims '
                    'with a tolerance of 2 frames)')
            raw = engine.compute_pitch(plans['pitch'], pbatch)
            pseed = (engine.next_seed()
                     if ppost.delta_pitch_noise_stddev != 0 else 0)
            engine.process_pitch(
                ppost._post_opts(), raw, engine.RowLayout(batch=pbatch),
                seed=pseed, out=out[:, self.feat_dim:])
            self._keep_pitch = (pbatch, raw)
        self._last = (batch, base)     # keep device buffers alive
        return out, batch.frame_offsets, stats, utt_group

    # -- end-to-end: pinned host PCM in, pinned host features out ------------
    def run_host(self, host_pcm, starts, lengths, chunk_utts=512,
                 out_host=None):
        """Pipelines chunks of utterances over three streams

        `host_pcm` is a pinned int16 tensor, `starts`/`lengths` int64 arrays.
        Per-utterance CMVN only (speaker CMVN needs a global barrier).
        Returns (pinned float32 [total_frames, out_dim], frame_offsets).
        """
        torch = engine.require_cuda()
        if self.cmvn == 'speaker':
            raise NotImplementedError('run_host supports per-utterance cmvn')
        plans = self._plans()
        nutts = len(lengths)
        nframes = engine.num_frames_array(
            self.processor._frame_opts(), lengths)
        foffs = np.concatenate(([0], np.cumsum(nframes))).astype(np.int64)
        total = int(foffs[-1])
        if out_host is None:
            out_host = torch.empty((total, self.out_dim), dtype=torch.float32,
                                   pin_memory=True)
        chunks = [(b, min(b + chunk_utts, nutts))
                  for b in range(0, nutts, chunk_utts)]
        span = max(int(starts[e - 1] + lengths[e - 1] - starts[b]) + 64
                   for b, e in chunks)
        span = (span + 7) // 8 * 8
        max_rows = max(int(foffs[e] - foffs[b]) for b, e in chunks)
        nslots = 3
        # streams and slot buffers persist across calls: PyTorch's caching
        # allocator pools are per stream, fresh streams would re-cudaMalloc
        # (and later cudaFree, a device-wide sync) every buffer on every call
        state = getattr(self, '_host_state', None)
        if (state is None or state['span'] < span
                or state['rows'] < max_rows):
            state = {
                'streams': tuple(torch.cuda.Stream() for _ in range(3)),
                'span': span, 'rows': max_rows,
                'pcm': [torch.empty(span, dtype=torch.int16, device='cuda')
                        for _ in range(nslots)],
                'out': [torch.empty((max_rows, self.out_dim),
                                    dtype=torch.float32, device='cuda')
                        for _ in range(nslots)],
                'base': [torch.empty((max_rows, self.base_dim),
                                     dtype=torch.float32, device='cuda')
                         for _ in range(nslots)]}
            self._host_state = state
        s_in, s_c, s_out = state['streams']
        pcm_slots, out_slots = state['pcm'], state['out']
        base_slots = state['base']
        for s in (s_in, s_c, s_out):      # order after the caller's stream
            s.wait_stream(torch.cuda.current_stream())
        free_ev = [None] * nslots        # D2H of the slot's previous use
        keep = []
        for i, (b, e) in enumerate(chunks):
            slot = i % nslots
            begin = int(starts[b])
            n = int(starts[e - 1] + lengths[e - 1]) - begin
            with torch.cuda.stream(s_in):
                # the (small) batch descriptor goes first: queued behind the
                # PCM copies of later chunks it would hold back this chunk's
                # kernels and drain the slot ring
                packed = engine.PackedAudio.from_packed(
                    None, starts[b:e] - begin, lengths[b:e],
                    dev=pcm_slots[slot])
                batch = engine.Batch(plans['feat'], packed)
                if free_ev[slot] is not None:
                    s_in.wait_event(free_ev[slot])
                pcm_slots[slot][:n].copy_(
                    host_pcm[begin:begin + n], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            rows = int(foffs[e] - foffs[b])
            with torch.cuda.stream(s_c):
                s_c.wait_event(ev_in)
                out_dev = out_slots[slot][:rows]
                self.run_device(packed, out=out_dev, plans=plans,
                                base_buf=base_slots[slot], batch=batch)
                keep.append((packed, self._last))
                ev_c = torch.cuda.Event()
                ev_c.record(s_c)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c)
                out_host[int(foffs[b]):int(foffs[e])].copy_(
                    out_dev, non_blocking=True)
                ev_out = torch.cuda.Event()
                ev_out.record(s_out)
                free_ev[slot] = ev_out
        for s in (s_in, s_c, s_out):
            s.synchronize()
        return out_host, foffs
