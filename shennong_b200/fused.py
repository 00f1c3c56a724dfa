"""Device-resident fused pipeline: features -> (VAD) -> CMVN -> delta ‖ pitch

This is the batched form of the reference's two-pass schedule
(shennong/pipeline.py:541-567, 570-648): everything between the packed int16
PCM and the final ``[frames, D]`` matrix stays in HBM and runs as a short
sequence of ``libsnb`` launches on one stream:

  pass one (per batch / chunk)
    snb_compute_features        base features            [F, d]
    snb_compute_features (energy) + snb_vad_energy        (with_vad only)
    snb_cmvn_accumulate         per-utterance f64 stats   [U, 2, d+1]
    snb_compute_pitch + snb_process_pitch   written into the last columns
  pass two (per batch, or per block of whole speakers)
    snb_cmvn_reduce_groups      per-speaker stats         (by_speaker only)
    snb_cmvn_norm_from_stats    f32 (offset, scale)
    snb_cmvn_apply_deltas       normalise + deltas, written with the final
                                leading dimension (column block 0)

so the concatenation of the reference (features.py:350-437) is a strided
write, not a copy.  ``run_device`` runs both passes on one resident batch;
:mod:`shennong_b200.stream` pipelines chunks of a host corpus through them
(``run_host`` is the pre-packed pinned-buffer entry used by the benchmark).
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
        self.simple = cmvn is None and delta is None
        self.valid_rows = None
        self._runner = None

    # -- plans ---------------------------------------------------------------
    def _plans(self):
        p = self.processor
        plans = {'feat': engine.feature_plan(
            p._frame_opts(), p._mel_opts(), p._feat_opts())}
        if self.vad is not None and self.cmvn is not None:
            e = self.energy
            plans['energy'] = engine.feature_plan(
                e._frame_opts(), None, e._feat_opts())
        if self.pitch is not None:
            plans['pitch'] = engine.pitch_plan(self.pitch[0]._pitch_opts())
        return plans

    def make_batches(self, plans, packed, warps=None):
        """The ragged-batch descriptors of every plan for `packed` (small
        uploads queued on the current stream)"""
        return {name: engine.Batch(plan, packed,
                                   warps if name == 'feat' else None)
                for name, plan in plans.items()}

    # -- pass one --------------------------------------------------------------
    def pass_one(self, packed, batches, plans, seed, base, out, stats=None):
        """Base features into `base` [F, base_dim] (may be a column view of
        `out`), per-utterance CMVN statistics into `stats` (allocated when
        None and CMVN is on), pitch columns into ``out[:, feat_dim:]``.
        Returns `stats`."""
        p = self.processor
        batch = batches['feat']
        layout = engine.RowLayout(batch=batch)
        fseed = seed if p.dither != 0 else 0
        engine.compute_features(plans['feat'], batch, seed=fseed, out=base)
        if self.cmvn is not None:
            weights = None
            if self.vad is not None:
                ebatch = batches['energy']
                eseed = (seed ^ 0x5851f42d4c957f2d
                         if self.energy.dither != 0 else 0)
                e64 = engine.compute_features(
                    plans['energy'], ebatch, seed=eseed, float64=True)
                e32 = engine.f64_to_f32(e64)
                v = self.vad
                weights = engine.vad_energy(
                    e32, engine.RowLayout(batch=ebatch), v.energy_threshold,
                    v.energy_mean_scale, v.frames_context,
                    v.proportion_threshold)
            stats = engine.cmvn_accumulate(base, layout, weights, out=stats)
        self.valid_rows = None
        if self.pitch is not None:
            pproc, ppost = self.pitch
            pbatch = batches['pitch']
            frames = np.diff(batch.frame_offsets)
            pframes = np.diff(pbatch.frame_offsets)
            same = np.array_equal(frames, pframes)
            if not same:
                diff = np.abs(frames - pframes)
                if diff.max() > 2:
                    u = int(np.argmax(diff))
                    raise ValueError(
                        'features differs number of frames, and greater than '
                        'tolerance: |{} - {}| > 2'.format(
                            frames[u], pframes[u]))
                # the reference trims the longer of the two
                # (Features.concatenate(tolerance=2), pipeline.py:639-641)
                self.valid_rows = np.minimum(frames, pframes)
            raw = engine.compute_pitch(plans['pitch'], pbatch)
            pseed = (seed ^ 0x14057b7ef767814f
                     if ppost.delta_pitch_noise_stddev != 0 else 0)
            engine.process_pitch(
                ppost._post_opts(), raw, engine.RowLayout(batch=pbatch),
                seed=pseed, out=out[:, self.feat_dim:],
                out_layout=None if same else layout)
        self._last = (batches, base)
        return stats

    # -- pass two --------------------------------------------------------------
    def pass_two(self, base, out, frame_offsets, ustats, group=None,
                 ngroups=0, layout=None, norm_out=None):
        """Normalisation table from the statistics and the normalise + delta
        launch over `base` rows described by `frame_offsets` (host int64
        [U+1]) or `layout`.  `group` (host int array) pools the
        per-utterance statistics by speaker.  Returns the statistics the
        normalisation used ([U or ngroups, 2, d+1] float64 device tensor)."""
        torch = engine.require_cuda()
        if layout is None:
            layout = engine.RowLayout(frame_offsets=frame_offsets)
        stats, utt_group, norm = ustats, None, None
        if self.cmvn is not None:
            if group is not None:
                group = np.asarray(group, dtype=np.int64)
                order = np.argsort(group, kind='stable')
                ptr = np.concatenate(
                    ([0], np.cumsum(np.bincount(group, minlength=ngroups))))
                stats = engine.cmvn_reduce_groups(ustats, ptr, order, ngroups)
                utt_group = torch.from_numpy(
                    group.astype(np.int32)).to('cuda', non_blocking=True)
            norm = engine.cmvn_norm(stats, self.norm_vars, False,
                                    out=norm_out)
        order = self.delta.order if self.delta is not None else 0
        window = self.delta.window if self.delta is not None else 1
        engine.deltas(base, layout, order, window, norm=norm,
                      utt_group=utt_group, out=out[:, :self.feat_dim])
        self._last = (layout, norm, utt_group, stats)
        return stats

    def run_device(self, packed, speakers=None, warps=None, seed=None,
                   out=None, plans=None, base_buf=None, batch=None,
                   batches=None, stats_out=None, norm_out=None):
        """Runs the whole pipeline on a PackedAudio already on the device

        Returns (out [total_frames, out_dim] device tensor, frame_offsets
        int64 [U+1] host array, stats float64 device tensor or None,
        group index of each utterance or None).  All launches are queued on
        the current stream; nothing synchronises.  `base_buf` is an optional
        preallocated [>= total_frames, base_dim] float32 buffer for the
        intermediate base features (chunked pipelines reuse it); `batches`
        the descriptors of :meth:`make_batches` already created for `packed`;
        `stats_out` an optional [U, 2, base_dim + 1] float64 device tensor
        for the per-utterance statistics, `norm_out` one [U or groups, 2,
        base_dim] float32 for the normalisation table (both are what a peer
        needs to redo pass two on the base rows:
        :class:`shennong_b200.distributed.ChunkCollector`).  When features and pitch disagree
        by one or two frames on some utterance, ``self.valid_rows`` holds the
        rows to keep per utterance (None otherwise).
        """
        torch = engine.require_cuda()
        plans = plans or self._plans()
        if batches is None:
            batches = self.make_batches(plans, packed, warps)
            if batch is not None:
                batches['feat'] = batch
        batch = batches['feat']
        if seed is None:
            seed = engine.next_seed()
        total = batch.total_frames
        if out is None:
            out = torch.empty((total, self.out_dim), dtype=torch.float32,
                              device='cuda')
        if self.simple:
            base = out[:, :self.base_dim]
        elif base_buf is not None:
            base = base_buf[:total]
        else:
            base = torch.empty((total, self.base_dim), dtype=torch.float32,
                               device='cuda')
        ustats = self.pass_one(packed, batches, plans, seed, base, out,
                               stats=stats_out)
        keep = self._last
        stats, group = None, None
        if not self.simple:
            ngroups = 0
            if self.cmvn == 'speaker':
                if speakers is None:
                    raise ValueError('speakers are required for cmvn by '
                                     'speaker')
                names = sorted(set(speakers))
                index = {s: i for i, s in enumerate(names)}
                group = np.array([index[s] for s in speakers], dtype=np.int64)
                ngroups = len(names)
                self._group_names = names
            stats = self.pass_two(
                base, out, None, ustats, group, ngroups,
                layout=engine.RowLayout(batch=batch), norm_out=norm_out)
            group = self._last[2]
        self._last = (batch, base, keep, self._last)   # keep buffers alive
        return out, batch.frame_offsets, stats, group

    # -- end-to-end: pinned host PCM in, pinned host features out ------------
    def run_host(self, host_pcm, starts, lengths, chunk_utts=None,
                 out_host=None, speakers=None):
        """Streams chunks of a packed pinned PCM buffer through the pipeline
        (H2D / compute / D2H on three streams, :mod:`shennong_b200.stream`)

        `host_pcm` is a pinned int16 tensor, `starts`/`lengths` int64 arrays
        (utterances in buffer order; ordered by speaker for CMVN by speaker).
        Returns (pinned float32 [total_frames, out_dim], frame_offsets).
        """
        from shennong_b200 import stream
        if self._runner is None or (
                chunk_utts and self._runner.chunk_utts != chunk_utts):
            self._runner = stream.StreamRunner(self, chunk_utts=chunk_utts)
        source = stream.PackedSource(host_pcm, starts, lengths)
        out, plan, stats = self._runner.run(
            source, speakers=speakers, out_host=out_host)
        self.host_stats = stats
        return out, plan.foffs
