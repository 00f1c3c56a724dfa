"""Streamed, bounded-memory execution of a :class:`FusedPipeline` over a corpus

This is the back end of the batch entry points of the host API
(``FeaturesProcessor.process_all``, ``pipeline.extract_features``; in the
reference: a joblib pool over utterances, shennong/processor/base.py:97-107
and shennong/pipeline.py:541-567).  The corpus never sits in device memory as
a whole: utterances travel in *chunks* through three CUDA streams

    host threads: load / pack PCM into pinned staging   (AudioSource)
    s_in   : H2D of the chunk's int16 PCM (+ its batch descriptors)
    s_c    : the fused pipeline on the chunk (libsnb launches)
    s_out  : D2H of the finished rows into the pinned result

with a ring of `nslots` device slots, so that device memory is
O(nslots x chunk) whatever the corpus.  CMVN by speaker needs every frame of
a speaker before the first one can be normalised -- the reference's two-pass
barrier (pipeline.py:543-557).  Here the utterances are ordered by speaker and
grouped in *blocks* of whole speakers whose base features fit a byte budget:
pass 1 streams the block's chunks (features, VAD weights, per-utterance
float64 statistics, pitch columns), one deterministic reduction gives the
speaker statistics, pass 2 is ONE normalise + delta launch over the block, and
the block's rows leave through s_out while the next block's pass 1 runs.

Multi-GPU (one process per GPU, torch.distributed): every rank streams its own
shard; with ``gather`` each finished piece is all-gathered over NCCL on a
fourth stream while the next piece is being computed, and every rank copies
the rows of all ranks to its host result (collection time).
"""

import concurrent.futures
import os
import threading

import numpy as np

from shennong_b200 import _lib, engine
from shennong_b200.audio import Audio

_ALIGN = 8


# --------------------------------------------------------------------------
# PCM sources
# --------------------------------------------------------------------------
class PackedSource:
    """Utterances already packed in ONE pinned int16 tensor, in processing
    order, every start on a multiple of 8 samples (bench, device tests)"""
    pinned = True

    def __init__(self, host, starts, lengths):
        self.host = host
        self.starts = np.ascontiguousarray(starts, dtype=np.int64)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        self.nutts = len(self.lengths)

    def span(self, b, e):
        return int(self.starts[e - 1] + self.lengths[e - 1]
                   - self.starts[b]) + 64

    def window(self, b, e, staging=None):
        begin = int(self.starts[b])
        n = int(self.starts[e - 1] + self.lengths[e - 1]) - begin
        return self.host[begin:begin + n], self.starts[b:e] - begin


class AudioSource:
    """Utterances as items loaded on demand into pinned staging buffers

    An item is a numpy array / Audio (converted to int16 with the
    reference's scaling, shennong/audio.py:495-518) or a
    ``(path, data_offset, first_sample, nsamples)`` tuple naming a segment of
    a mono 16-bit PCM WAV file: its bytes are read by ``readinto`` straight
    into the pinned buffer the DMA engine reads from (no intermediate numpy
    array, no per-utterance Python object).  `lengths` are the sample counts.
    """
    pinned = False

    def __init__(self, items, lengths, workers=8):
        self.items = items
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        self.nutts = len(self.lengths)
        padded = (self.lengths + _ALIGN - 1) // _ALIGN * _ALIGN
        self.pstarts = np.concatenate(([0], np.cumsum(padded)))
        self.workers = max(1, int(workers))

    def span(self, b, e):
        return int(self.pstarts[e] - self.pstarts[b]) + 64

    def _load(self, i, view):
        item = self.items[i]
        n = int(self.lengths[i])
        if isinstance(item, tuple):
            path, offset, first, _ = item
            with open(path, 'rb', buffering=0) as fh:
                fh.seek(offset + 2 * first)
                got = fh.readinto(memoryview(view[:n]).cast('B'))
            if got != 2 * n:
                raise ValueError(f'{path}: cannot read file, truncated data')
            return
        data = item.data if hasattr(item, 'sample_rate') else item
        if data.dtype != np.int16:
            data = _to_int16(data)
        view[:n] = data

    def window(self, b, e, staging):
        """Packs utterances [b, e) into `staging` (pinned int16 tensor);
        returns (tensor slice, relative starts)"""
        rel = self.pstarts[b:e] - self.pstarts[b]
        files = [i for i in range(b, e) if isinstance(self.items[i], tuple)]
        if files:
            # WAV payloads: one native call, `workers` threads preading into
            # the staging buffer (snb_read_segments, no GIL)
            self._read_files(files, rel, b, staging.data_ptr())
        if len(files) < e - b:
            view = staging.numpy()
            for i in range(b, e):
                if not isinstance(self.items[i], tuple):
                    s = int(rel[i - b])
                    self._load(i, view[s:s + int(self.lengths[i])])
        n = int(rel[-1] + self.lengths[e - 1])
        return staging[:n], rel

    def _read_files(self, files, rel, b, base_ptr):
        import ctypes
        n = len(files)
        paths = (ctypes.c_char_p * n)()
        offsets = np.empty(n, dtype=np.int64)
        nbytes = np.empty(n, dtype=np.int64)
        dst = np.empty(n, dtype=np.uint64)
        for j, i in enumerate(files):
            path, offset, first, _ = self.items[i]
            paths[j] = os.fsencode(path)
            offsets[j] = offset + 2 * first
            nbytes[j] = 2 * int(self.lengths[i])
            dst[j] = base_ptr + 2 * int(rel[i - b])
        failed = ctypes.c_int64(-1)
        code = _lib.lib().snb_read_segments(
            paths, offsets.ctypes.data_as(ctypes.c_void_p),
            nbytes.ctypes.data_as(ctypes.c_void_p),
            dst.ctypes.data_as(ctypes.c_void_p), n, self.workers,
            ctypes.byref(failed))
        if code != 0:
            bad = self.items[files[failed.value]][0] if failed.value >= 0 \
                else '?'
            raise ValueError(f'{bad}: cannot read file, truncated data')


def _to_int16(data):
    """Audio.astype(np.int16) of the reference (audio.py:495-518)"""
    if data.dtype == np.int32:
        return (data / 2**15).astype(np.int16)
    return (data * 2**15).astype(np.int16)


# --------------------------------------------------------------------------
# plan: chunks, blocks, rows
# --------------------------------------------------------------------------
class StreamPlan:
    """Row geometry and chunking of one rank's utterances (processing order)

    Attributes: ``frames`` rows of every utterance in the result matrix,
    ``valid`` rows that hold data (features and pitch may differ by up to two
    frames: the reference trims the longer one, features.py:350-437),
    ``foffs`` row offsets, ``chunks`` [(b, e)] utterance ranges, ``blocks``
    [(first chunk, last chunk + 1)] -- one chunk per block unless CMVN is by
    speaker.
    """

    def __init__(self, pipe, lengths, groups=None, chunk_utts=512,
                 chunk_samples=96_000_000, block_bytes=8 << 30):
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        self.nutts = n = len(lengths)
        self.frames = engine.num_frames_array(
            pipe.processor._frame_opts(), lengths).astype(np.int64)
        self.valid = self.frames
        if pipe.pitch is not None:
            pframes = engine.pitch_num_frames_array(
                pipe.pitch[0]._pitch_opts(), lengths)
            diff = np.abs(pframes - self.frames)
            if n and diff.max() > 2:
                u = int(np.argmax(diff))
                raise ValueError(
                    'features differs number of frames, and greater than '
                    'tolerance: |{} - {}| > 2'.format(
                        self.frames[u], pframes[u]))
            self.valid = np.minimum(self.frames, pframes)
        self.foffs = np.concatenate(([0], np.cumsum(self.frames)))
        self.total = int(self.foffs[-1])
        padded = (lengths + _ALIGN - 1) // _ALIGN * _ALIGN
        cum = np.concatenate(([0], np.cumsum(padded)))
        # blocks of whole groups (speakers) within the byte budget
        row_bytes = 4 * (pipe.base_dim + pipe.out_dim)
        if groups is None:
            block_ends = None
        else:
            groups = np.asarray(groups)
            change = np.flatnonzero(groups[1:] != groups[:-1]) + 1
            unit_ends = np.append(change, n)
            budget = max(1, block_bytes // row_bytes)
            block_ends, start_row = [], 0
            for k, end in enumerate(unit_ends):
                nxt = unit_ends[k + 1] if k + 1 < len(unit_ends) else None
                if nxt is None or self.foffs[nxt] - start_row > budget:
                    block_ends.append(int(end))
                    start_row = self.foffs[end]
        self.chunks, self.blocks = [], []
        b = 0
        limits = block_ends if block_ends is not None else [n]
        for limit in limits:
            first = len(self.chunks)
            while b < limit:
                e = int(np.searchsorted(cum, cum[b] + chunk_samples, 'right')) - 1
                e = max(b + 1, min(e, b + chunk_utts, limit))
                self.chunks.append((b, e))
                b = e
            if block_ends is not None and len(self.chunks) > first:
                self.blocks.append((first, len(self.chunks)))
        if block_ends is None:
            self.blocks = [(i, i + 1) for i in range(len(self.chunks))]

    def chunk_rows(self, c):
        b, e = self.chunks[c]
        return int(self.foffs[b]), int(self.foffs[e])

    @property
    def max_chunk_rows(self):
        return max((self.chunk_rows(c)[1] - self.chunk_rows(c)[0]
                    for c in range(len(self.chunks))), default=0)

    @property
    def max_block_rows(self):
        return max((self.chunk_rows(l - 1)[1] - self.chunk_rows(f)[0]
                    for f, l in self.blocks), default=0)


# --------------------------------------------------------------------------
# collectors: where finished rows go
# --------------------------------------------------------------------------
class HostCollector:
    """D2H of every finished piece into a pinned [total_rows, D] result"""

    def __init__(self, out_host, stream):
        self.out_host = out_host
        self.stream = stream
        self.pieces = 0

    def put(self, dev_rows, row0, ready):
        """queues the copy of `dev_rows` to rows [row0, ...); returns the
        event after which `dev_rows` may be overwritten"""
        torch = engine._torch()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.out_host[row0:row0 + dev_rows.shape[0]].copy_(
                dev_rows, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.pieces += 1
        return done

    def finish(self):
        self.stream.synchronize()


class _NoStream:
    """stands for a CUDA stream when the collector runs on CPU tensors
    (gloo tests of the collection logic)"""
    def wait_event(self, event):
        pass

    def synchronize(self):
        pass


class GatherCollector:
    """All-gather of every finished piece over NCCL (on its own stream, so
    piece k travels while piece k + 1 is computed), then D2H of the rows of
    ALL ranks.  `plans` are the StreamPlans of every rank (each rank derives
    all of them from the corpus index: no negotiation at run time); the host
    result is rank-major: rows of rank r start at ``rank_row0[r]``.

    With `stream` and `comm_stream` None the collector works on CPU tensors
    (gloo): the same bookkeeping, no streams.
    """

    def __init__(self, out_host, stream, comm_stream, plans, rank, dim,
                 group=None):
        torch = engine._torch()
        self.cuda = stream is not None
        self.out_host = out_host
        self.stream = stream if self.cuda else _NoStream()
        self.comm = comm_stream if self.cuda else _NoStream()
        self.plans, self.rank, self.size = plans, rank, len(plans)
        self.group = group
        self.rounds = max(len(p.chunks) for p in plans)
        self.rank_row0 = np.concatenate(
            ([0], np.cumsum([p.total for p in plans])))
        self.height = max(max(p.max_chunk_rows for p in plans), 1)
        self.dim = dim
        self.device = 'cuda' if self.cuda else 'cpu'
        self.gbuf = [torch.empty((self.size, self.height, dim),
                                 dtype=torch.float32, device=self.device)
                     for _ in range(2)]
        self.gfree = [None, None]
        self.dummy = torch.zeros((self.height, dim), dtype=torch.float32,
                                 device=self.device)
        self.sendbuf = None
        self.pieces = 0

    def _on(self, stream):
        import contextlib
        if not self.cuda:
            return contextlib.nullcontext()
        return engine._torch().cuda.stream(stream)

    def _event(self, stream):
        if not self.cuda:
            return None
        event = engine._torch().cuda.Event()
        event.record(stream)
        return event

    def _rows(self, r, piece):
        plan = self.plans[r]
        if piece >= len(plan.chunks):
            return 0, 0
        a, b = plan.chunk_rows(piece)
        return a, b - a

    def put(self, dev_rows, row0, ready, padded=None):
        """`padded` is the slot tensor `dev_rows` is a prefix of (the round's
        height is read from it; rows beyond the piece are ignored by the
        receivers).  Returns the event after which the source may be
        overwritten."""
        import torch.distributed as dist
        torch = engine._torch()
        piece = self.pieces
        self.pieces += 1
        h = max(max(self._rows(r, piece)[1] for r in range(self.size)), 1)
        g = piece % 2
        with self._on(self.comm):
            if ready is not None:
                self.comm.wait_event(ready)
            if self.gfree[g] is not None:
                self.comm.wait_event(self.gfree[g])
            if padded is not None and padded.shape[0] >= h:
                send = padded[:h]
            elif dev_rows.shape[0] == h:
                send = dev_rows
            elif dev_rows.shape[0] == 0:
                send = self.dummy[:h]
            else:                      # a piece shorter than the round's height
                if self.sendbuf is None:
                    self.sendbuf = torch.empty_like(self.dummy)
                self.sendbuf[:dev_rows.shape[0]].copy_(dev_rows)
                send = self.sendbuf[:h]
            # the first size * h rows of the (flat) receive buffer
            got = self.gbuf[g].view(-1)[:self.size * h * self.dim].view(
                self.size, h, self.dim)
            dist.all_gather_into_tensor(
                got.view(self.size * h, self.dim), send.contiguous(),
                group=self.group)
            sent = self._event(self.comm)
        with self._on(self.stream):
            if sent is not None:
                self.stream.wait_event(sent)
            for r in range(self.size):
                a, n = self._rows(r, piece)
                if n:
                    base = int(self.rank_row0[r])
                    self.out_host[base + a:base + a + n].copy_(
                        got[r, :n], non_blocking=True)
            self.gfree[g] = self._event(self.stream)
        return sent

    def finish(self):
        while self.pieces < self.rounds:      # ranks with fewer pieces
            self.put(self.dummy[:0], 0, None)
        self.comm.synchronize()
        self.stream.synchronize()


# --------------------------------------------------------------------------
# the runner
# --------------------------------------------------------------------------
class StreamRunner:
    """Runs `pipe` (a FusedPipeline) over a PCM source in bounded memory

    Buffers, streams and plans persist across calls (PyTorch's caching
    allocator pools are per stream: fresh streams would cudaMalloc -- and
    later cudaFree, a device-wide synchronisation -- on every call).
    """

    # Streams and slot buffers are shared by every runner of the process (the
    # batch entry points build a runner per call): keyed by what they hold,
    # they only grow.  `_pool_lock` serialises the runs that use them.
    _pool = {}
    _pool_lock = threading.Lock()

    def __init__(self, pipe, chunk_utts=None, chunk_samples=None,
                 block_bytes=None, nslots=3):
        self.pipe = pipe
        # defaults: 512 utterances or 96 M samples (192 MB of PCM) per chunk,
        # 8 GiB of base + final features per block of speakers
        env = os.environ.get
        default_utts, default_samples = 512, 96_000_000
        if pipe.pitch is not None:
            # the pitch tracker follows one utterance per warp, sequentially
            # over its frames: a chunk must fill a whole round of it (4 736
            # utterances on a B200) or the kernel runs mostly empty
            wave = int(_lib.lib().snb_pitch_wave_utts(
                engine.pitch_plan(pipe.pitch[0]._pitch_opts()).handle))
            default_utts = max(default_utts, wave)
            default_samples = max(default_samples, wave * 170_000)
        self.chunk_utts = int(
            chunk_utts or env('SNB_STREAM_CHUNK_UTTS', default_utts))
        self.chunk_samples = int(
            chunk_samples or env('SNB_STREAM_CHUNK_SAMPLES', default_samples))
        self.block_bytes = int(
            block_bytes or env('SNB_STREAM_BLOCK_BYTES', 8 << 30))
        self.nslots = int(nslots)
        self._state = None
        self._lock = threading.Lock()

    def plan(self, lengths, groups=None):
        return StreamPlan(self.pipe, lengths, groups, self.chunk_utts,
                          self.chunk_samples, self.block_bytes)

    # -- buffers ---------------------------------------------------------------
    def _buffers(self, span, rows, block_rows, staging):
        torch = engine.require_cuda()
        pipe = self.pipe
        key = (torch.cuda.current_device(), self.nslots, pipe.out_dim,
               pipe.base_dim)
        st = StreamRunner._pool.get(key)
        if st is None:
            st = StreamRunner._pool[key] = {
                'streams': tuple(torch.cuda.Stream() for _ in range(4)),
                'span': 0, 'rows': 0, 'block_rows': 0, 'staging': 0}
        self._state = st
        if st['span'] < span:
            st['span'] = span
            st['pcm'] = [torch.empty(span, dtype=torch.int16, device='cuda')
                         for _ in range(self.nslots)]
        if staging and st['staging'] < span:
            st['staging'] = span
            st['stage'] = [torch.empty(span, dtype=torch.int16,
                                       pin_memory=True)
                           for _ in range(self.nslots)]
        by_block = pipe.cmvn == 'speaker'
        if not by_block and st['rows'] < rows:
            st['rows'] = rows
            st['out'] = [torch.empty((rows, pipe.out_dim),
                                     dtype=torch.float32, device='cuda')
                         for _ in range(self.nslots)]
            st['base'] = [torch.empty((rows, pipe.base_dim),
                                      dtype=torch.float32, device='cuda')
                          for _ in range(self.nslots)]
        if by_block and st['block_rows'] < block_rows:
            st['block_rows'] = block_rows
            st['bout'] = [torch.empty((block_rows, pipe.out_dim),
                                      dtype=torch.float32, device='cuda')
                          for _ in range(2)]
            st['bbase'] = [torch.empty((block_rows, pipe.base_dim),
                                       dtype=torch.float32, device='cuda')
                           for _ in range(2)]
        return st

    # -- run ---------------------------------------------------------------------
    def run(self, source, plan=None, speakers=None, warps=None, out_host=None,
            gather=None, seed=None):
        """Extracts the whole source

        Returns (pinned float32 [rows, out_dim] host tensor, plan, stats):
        `stats` is a float64 numpy array [nutts or ngroups, 2, base_dim + 1]
        (None without CMVN) and, by speaker, comes with ``self.group_names``
        / ``self.utt_group``.  With a `collector` the rows go where it sends
        them and the first element is its ``out_host``.  `gather` is
        ``(plans of every rank, this rank[, process group])``: the pieces are
        all-gathered (:class:`GatherCollector`, kept as ``self.collector``)
        and the result holds the rows of all ranks, rank-major.
        """
        torch = engine.require_cuda()
        pipe = self.pipe
        by_speaker = pipe.cmvn == 'speaker'
        groups = None
        if by_speaker:
            if speakers is None:
                raise ValueError('speakers are required for cmvn by speaker')
            names, groups = np.unique(np.asarray(speakers), return_inverse=True)
            if np.any(np.diff(groups) < 0):
                raise ValueError('utterances must be ordered by speaker')
            self.group_names = [str(s) for s in names]
            self.utt_group = groups
        if plan is None:
            plan = self.plan(source.lengths, groups)
        with StreamRunner._pool_lock:
            return self._run(torch, source, plan, groups, warps, out_host,
                             gather, seed)

    def _run(self, torch, source, plan, groups, warps, out_host, gather,
             seed):
        pipe = self.pipe
        plans = pipe._plans()
        by_speaker = pipe.cmvn == 'speaker'
        span = max((source.span(b, e) for b, e in plan.chunks), default=64)
        span = (span + 7) // 8 * 8
        st = self._buffers(span, plan.max_chunk_rows, plan.max_block_rows,
                           not source.pinned)
        s_in, s_c, s_out, s_comm = st['streams']
        cur = torch.cuda.current_stream()
        for s in (s_in, s_c, s_out, s_comm):
            s.wait_stream(cur)
        if gather is not None:
            gplans, grank = gather[0], gather[1]
            if out_host is None:
                out_host = torch.empty(
                    (sum(p.total for p in gplans), pipe.out_dim),
                    dtype=torch.float32, pin_memory=True)
            collector = GatherCollector(
                out_host, s_out, s_comm, gplans, grank, pipe.out_dim,
                gather[2] if len(gather) > 2 else None)
        else:
            if out_host is None:
                out_host = torch.empty((plan.total, pipe.out_dim),
                                       dtype=torch.float32, pin_memory=True)
            collector = HostCollector(out_host, s_out)
        self.collector = collector
        stats_dev = None
        if pipe.cmvn is not None:
            ng = len(self.group_names) if by_speaker else plan.nutts
            with torch.cuda.stream(s_c):
                stats_dev = torch.empty((ng, 2, pipe.base_dim + 1),
                                        dtype=torch.float64, device='cuda')
        nchunks = len(plan.chunks)
        # host-side packing runs ahead of the device in a worker thread
        loader = None
        if not source.pinned:
            loader = concurrent.futures.ThreadPoolExecutor(1)
        stage_free = [None] * self.nslots     # H2D of the staging slot done
        pcm_free = [None] * self.nslots       # kernels reading the PCM slot done
        out_free = [None] * self.nslots       # D2H / gather of the out slot done
        if seed is None:
            seed = engine.next_seed()

        def load(c):
            slot = c % self.nslots
            b, e = plan.chunks[c]
            if source.pinned:
                return source.window(b, e)
            if stage_free[slot] is not None:
                stage_free[slot].synchronize()
            return source.window(b, e, st['stage'][slot])

        pending = None
        if loader is not None and nchunks:
            pending = loader.submit(load, 0)

        def upload(c):
            """H2D of chunk c: returns (packed, batches, event)"""
            nonlocal pending
            slot = c % self.nslots
            b, e = plan.chunks[c]
            if loader is not None:
                host, rel = pending.result()
                pending = loader.submit(load, c + 1) if c + 1 < nchunks else None
            else:
                host, rel = load(c)
            with torch.cuda.stream(s_in):
                # the (small) batch descriptors go first: queued behind the
                # PCM copies of later chunks they would hold back this
                # chunk's kernels and drain the slot ring
                packed = engine.PackedAudio.from_packed(
                    None, rel, source.lengths[b:e], dev=st['pcm'][slot])
                batches = pipe.make_batches(
                    plans, packed, None if warps is None else warps[b:e])
                if pcm_free[slot] is not None:
                    s_in.wait_event(pcm_free[slot])
                st['pcm'][slot][:host.numel()].copy_(host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
                if not source.pinned:
                    stage_free[slot] = ev
            return packed, batches, ev

        if not by_speaker:
            for c in range(nchunks):
                slot = c % self.nslots
                b, e = plan.chunks[c]
                r0, r1 = plan.chunk_rows(c)
                packed, batches, ev_in = upload(c)
                with torch.cuda.stream(s_c):
                    s_c.wait_event(ev_in)
                    if out_free[slot] is not None:
                        s_c.wait_event(out_free[slot])
                    out_dev = st['out'][slot][:r1 - r0]
                    pipe.run_device(
                        packed, out=out_dev, plans=plans, seed=seed + c,
                        base_buf=st['base'][slot], batches=batches,
                        stats_out=None if stats_dev is None
                        else stats_dev[b:e])
                    ev_c = torch.cuda.Event()
                    ev_c.record(s_c)
                    pcm_free[slot] = ev_c
                if isinstance(collector, GatherCollector):
                    out_free[slot] = collector.put(
                        out_dev, r0, ev_c, padded=st['out'][slot])
                else:
                    out_free[slot] = collector.put(out_dev, r0, ev_c)
        else:
            block_free = [None, None]
            for k, (first, last) in enumerate(plan.blocks):
                side = k % 2
                ub, ue = plan.chunks[first][0], plan.chunks[last - 1][1]
                row_a = plan.chunk_rows(first)[0]
                rows = plan.chunk_rows(last - 1)[1] - row_a
                bout, bbase = st['bout'][side][:rows], st['bbase'][side][:rows]
                with torch.cuda.stream(s_c):
                    if block_free[side] is not None:
                        s_c.wait_event(block_free[side])
                    ustats = torch.empty(
                        (ue - ub, 2, pipe.base_dim + 1), dtype=torch.float64,
                        device='cuda')
                for c in range(first, last):
                    slot = c % self.nslots
                    b, e = plan.chunks[c]
                    r0, r1 = plan.chunk_rows(c)
                    packed, batches, ev_in = upload(c)
                    with torch.cuda.stream(s_c):
                        s_c.wait_event(ev_in)
                        pipe.pass_one(
                            packed, batches, plans, seed + c,
                            base=bbase[r0 - row_a:r1 - row_a],
                            out=bout[r0 - row_a:r1 - row_a],
                            stats=ustats[b - ub:e - ub])
                        ev_c = torch.cuda.Event()
                        ev_c.record(s_c)
                        pcm_free[slot] = ev_c
                with torch.cuda.stream(s_c):
                    g = groups[ub:ue]
                    g0 = int(g[0])
                    gstats = pipe.pass_two(
                        bbase, bout, plan.foffs[ub:ue + 1] - row_a, ustats,
                        g - g0, int(g[-1]) - g0 + 1)
                    stats_dev[g0:int(g[-1]) + 1].copy_(gstats)
                    ev_b = torch.cuda.Event()
                    ev_b.record(s_c)
                done = None
                for c in range(first, last):
                    r0, r1 = plan.chunk_rows(c)
                    piece = bout[r0 - row_a:r1 - row_a]
                    done = collector.put(piece, r0, ev_b)
                block_free[side] = done
        stats = None
        if stats_dev is not None:
            with torch.cuda.stream(s_c):
                stats = stats_dev.cpu().numpy()
        s_c.synchronize()
        collector.finish()
        if loader is not None:
            loader.shutdown()
        return collector.out_host, plan, stats


# --------------------------------------------------------------------------
# corpus level: what the batch entry points of the host API call
# --------------------------------------------------------------------------
def wav_layouts(paths, nthreads=None):
    """[(data offset in bytes, nsamples, sample rate) or None] for `paths`:
    the header walk of every DISTINCT file on native threads
    (``snb_wav_scan_batch``); None for anything that is not a mono 16-bit PCM
    WAV file.  Nothing is cached: a file that changed between two calls is
    seen as it is now."""
    import ctypes
    distinct = list(dict.fromkeys(str(p) for p in paths))
    n = len(distinct)
    if n == 0:
        return []
    arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in distinct])
    offset = np.empty(n, dtype=np.int64)
    nsamples = np.empty(n, dtype=np.int64)
    rate = np.empty(n, dtype=np.int32)
    threads = nthreads or min(len(os.sched_getaffinity(0)), 32)
    _lib.check(_lib.lib().snb_wav_scan_batch(
        arr, n, offset.ctypes.data_as(ctypes.c_void_p),
        nsamples.ctypes.data_as(ctypes.c_void_p),
        rate.ctypes.data_as(ctypes.c_void_p), threads))
    found = {p: ((int(offset[i]), int(nsamples[i]), int(rate[i]))
                 if offset[i] >= 0 else None)
             for i, p in enumerate(distinct)}
    return [found[str(p)] for p in paths]


def audio_items(utts, sample_rate=None):
    """What :class:`AudioSource` reads for the utterances `utts`:
    (items, sample counts, all int16).  Segments of mono 16-bit PCM WAV files
    are described by (path, data offset, first sample, nsamples) and read
    straight into pinned memory; anything else is loaded as an Audio
    (utterances.py:171-176, audio.py:520-561 for the segment bounds).  With
    `sample_rate` the reference's process-time check of the signal's rate is
    applied (processor/base.py:415-419)."""
    items, lengths, int16 = [], [], True
    layouts = wav_layouts([utt.audio_file for utt in utts])
    for utt, layout in zip(utts, layouts):
        if layout is None:
            meta = Audio.scan(utt.audio_file)
            nchannels, rate = meta.nchannels, meta.sample_rate
        else:
            nchannels, rate = 1, layout[2]
        if nchannels != 1:
            raise ValueError(
                'signal must have one dimension, but it has {}'.format(
                    nchannels))
        if sample_rate is not None and rate != sample_rate:
            raise ValueError(
                'processor and signal mismatch in sample rates: '
                '{} != {}'.format(sample_rate, rate))
        if layout is None:
            audio = utt.load_audio()
            int16 = int16 and audio.dtype == np.int16
            items.append(audio)
            lengths.append(audio.nsamples)
            continue
        offset, nsamples = layout[0], layout[1]
        first, last = 0, nsamples
        if utt.tstart or utt.tstop:
            first = min(int(utt.tstart * rate), nsamples)
            last = max(first, min(int(utt.tstop * rate), nsamples))
        items.append((utt.audio_file, offset, first, last - first))
        lengths.append(last - first)
    return items, np.asarray(lengths, dtype=np.int64), int16


def extract_corpus(pipe, items, lengths, speakers=None, warps=None, njobs=1,
                   gather=True, runner=None):
    """Streams a corpus through `pipe`, sharded over the ranks of
    torch.distributed when it is initialised (one process per GPU)

    `speakers` (ordered: whole speakers contiguous) is only given for CMVN by
    speaker; shards then hold whole speakers and need no collective on the
    data path.  Returns (data, parts): `data` is the float32
    [rows, out_dim] host matrix and `parts` a list of (utterance indices,
    StreamPlan, first row in data, stats, group index per utterance) -- one
    entry, or one per rank when the rows were gathered.
    """
    from shennong_b200 import distributed
    runner = runner or StreamRunner(pipe)
    rank, world = distributed.world()
    by_speaker = pipe.cmvn == 'speaker'
    if world == 1:
        source = AudioSource(items, lengths, workers=njobs)
        out, plan, stats = runner.run(source, speakers=speakers, warps=warps)
        group = runner.utt_group if by_speaker else None
        return out.numpy(), [(np.arange(len(lengths)), plan, 0, stats, group)]
    frames = engine.num_frames_array(pipe.processor._frame_opts(), lengths)
    shards = distributed.shard_utterances(frames, world, speakers)
    mine = shards[rank]
    source = AudioSource([items[i] for i in mine], lengths[mine],
                         workers=njobs)

    def sub(seq, idx):
        return None if seq is None else [seq[i] for i in idx]
    plans = []
    for shard in shards:
        g = None
        if by_speaker:
            g = np.unique(np.asarray(sub(speakers, shard)),
                          return_inverse=True)[1]
        plans.append(runner.plan(lengths[shard], g))
    out, plan, stats = runner.run(
        source, plan=plans[rank], speakers=sub(speakers, mine),
        warps=None if warps is None else np.asarray(warps)[mine],
        gather=(plans, rank) if gather else None)
    group = runner.utt_group if by_speaker else None
    if not gather:
        return out.numpy(), [(mine, plan, 0, stats, group)]
    everything = distributed.all_gather_objects((stats, group))
    row0 = runner.collector.rank_row0
    return out.numpy(), [
        (shards[r], plans[r], int(row0[r]), everything[r][0],
         everything[r][1]) for r in range(world)]
