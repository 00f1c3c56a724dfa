"""Multi-GPU execution: shard utterances, gather features at collection time

The path shards naturally: utterances are independent units for
every stage except speaker-level CMVN.  One process per GPU (torchrun), each
rank extracts its own contiguous shard with the fused pipeline -- no
collective on the data path -- and the results are collected with one NCCL
all-gather of row counts followed by one all-gather of the (padded) feature
blocks over NVLink.  With ``cmvn by speaker`` the shards are speaker-aligned
(all the utterances of a speaker land on one rank) so that no statistics
all-reduce is needed; :func:`allreduce_stats` exists for callers that must
split a speaker.

Host-side logic is backend agnostic: the CPU tests run it on ``gloo`` with
world_size 2.
"""

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def world():
    """(rank, world_size) of the default process group, (0, 1) if none"""
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_utterances(costs, world_size, groups=None):
    """Splits utterances in `world_size` contiguous shards of balanced cost

    Parameters
    ----------
    costs : sequence of int
        Cost of each utterance (number of frames or samples)
    world_size : int
    groups : sequence, optional
        Group (speaker) of each utterance; when given, utterances are first
        ordered by group and a group is never split across shards

    Returns
    -------
    shards : list of int64 arrays
        Indices of the utterances of each rank; their concatenation is a
        permutation of ``range(len(costs))`` (identity when `groups` is None)
    """
    costs = np.asarray(costs, dtype=np.int64)
    n = len(costs)
    if world_size <= 0:
        raise ValueError('world_size must be strictly positive')
    if groups is None:
        order = np.arange(n, dtype=np.int64)
        unit_of = np.arange(n, dtype=np.int64)
    else:
        if len(groups) != n:
            raise ValueError('one group per utterance is expected')
        labels = {g: i for i, g in enumerate(sorted(set(groups), key=str))}
        ids = np.array([labels[g] for g in groups], dtype=np.int64)
        order = np.argsort(ids, kind='stable').astype(np.int64)
        unit_of = ids[order]
    # atomic units (utterances or whole groups) in order, with their cost
    boundaries = np.flatnonzero(np.diff(unit_of, prepend=-1))
    unit_cost = np.add.reduceat(costs[order], boundaries) if n else costs
    unit_end = np.append(boundaries[1:], n)
    total = int(unit_cost.sum()) if n else 0
    shards, start_unit, acc = [], 0, 0
    cum = np.cumsum(unit_cost) if n else np.zeros(0, np.int64)
    for rank in range(world_size):
        if rank == world_size - 1:
            stop_unit = len(unit_cost)
        else:
            target = total * (rank + 1) / world_size
            # first unit boundary whose cumulative cost reaches the target,
            # choosing the closer side
            stop_unit = int(np.searchsorted(cum, target, side='left'))
            if stop_unit < len(cum):
                before = cum[stop_unit - 1] if stop_unit > 0 else 0
                if cum[stop_unit] - target <= target - before:
                    stop_unit += 1
            stop_unit = max(stop_unit, start_unit)
            stop_unit = min(stop_unit, len(unit_cost))
        lo = boundaries[start_unit] if start_unit < len(boundaries) else n
        hi = unit_end[stop_unit - 1] if stop_unit > start_unit else lo
        shards.append(order[lo:hi])
        start_unit = stop_unit
        acc += 1
    return shards


def gather_rows(local, group=None):
    """All-gathers row blocks of different heights

    `local` is a [rows, D] tensor (CUDA with NCCL, CPU with gloo); returns
    (tensor [sum(rows), D] in rank order, int64 array of per-rank row counts).
    """
    import torch
    dist = _dist()
    rank, size = world()
    if size == 1:
        return local, np.array([local.shape[0]], dtype=np.int64)
    count = torch.tensor([local.shape[0]], dtype=torch.int64,
                         device=local.device)
    counts = torch.empty(size, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, count, group=group)
    counts_host = counts.cpu().numpy()
    height = int(counts_host.max())
    padded = local
    if local.shape[0] != height:
        padded = torch.zeros((height,) + tuple(local.shape[1:]),
                             dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    gathered = torch.empty((size * height,) + tuple(local.shape[1:]),
                           dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if int(counts_host.min()) == height:
        return gathered, counts_host      # equal shards: already contiguous
    blocks = [gathered[r * height:r * height + int(counts_host[r])]
              for r in range(size)]
    return torch.cat(blocks, dim=0), counts_host


def all_gather_objects(obj, group=None):
    """[obj of rank 0, ..., obj of rank n-1] on every rank (small host
    objects: CMVN statistics, speaker indices)"""
    dist = _dist()
    size = world()[1]
    if size == 1:
        return [obj]
    out = [None] * size
    dist.all_gather_object(out, obj, group=group)
    return out


class _DeviceBuffer:
    """CUDA array interface over a raw device pointer (zero-copy tensor)"""
    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {
            'shape': (int(nfloats),), 'typestr': '<f4',
            'data': (int(ptr), False), 'version': 2}


class PeerGather:
    """Collection by direct stores over NVLink

    Every rank owns a float32 result buffer of `nfloats`; the buffers are
    mapped in every process of the node through CUDA IPC
    (``snb_peer_buffer_*``) and :meth:`push` writes a block of finished rows
    at the same offset of ALL of them (``snb_gather_rows_ce`` /
    ``snb_gather_rows_bulk`` / ``snb_gather_rows``) on the current stream --
    the all-gather of the north star as posted writes, issued by the producer
    as soon as a chunk is done, with no rendezvous per chunk.  The pipeline
    can write its rows straight into the own buffer (:attr:`tensor`): the
    gather staging buffer IS the result.  :meth:`arrive` is the one
    synchronisation of a step: a 4-byte all-reduce queued behind the pushes.
    """

    def __init__(self, nfloats, group=None, fanout=1):
        """`fanout` > 1 (link-load experiments only): every peer buffer is
        `fanout` times as large and receives `fanout` copies of each push, so
        that two GPUs carry the NVLink traffic of ``1 + fanout`` ranks"""
        import ctypes
        import torch
        from shennong_b200 import _lib
        dist = _dist()
        self.rank, self.size = world()
        self.group = group
        self.nfloats = int(nfloats)
        self.fanout = max(1, int(fanout))
        L = _lib.lib()
        self._lib = L
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(L.snb_peer_buffer_create(
            self.nfloats * 4 * self.fanout, ctypes.byref(ptr), handle))
        self._own = ptr.value
        handles = all_gather_objects(bytes(handle), group)
        self._opened = []
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self._own)
                continue
            peer = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            _lib.check(L.snb_peer_buffer_open(buf, ctypes.byref(peer)))
            self._opened.append(peer.value)
            ptrs += [peer.value + 4 * self.nfloats * j
                     for j in range(self.fanout)]
        self._nptrs = len(ptrs)
        self._ptrs = (ctypes.c_void_p * self._nptrs)(*ptrs)
        self.tensor = torch.as_tensor(
            _DeviceBuffer(self._own, self.nfloats), device='cuda')
        self._flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        if self.size > 1:
            dist.barrier(group=group)

    def push(self, src, offset_floats, ctas=0, how='ce'):
        """`src` (contiguous float32 device tensor, a multiple of 4 elements)
        -> floats [offset, offset + src.numel()) of every rank's buffer

        `how`: 'ce' copy engines (one peer copy per rank on internal
        streams, joined to the current stream: no SM is used), 'bulk' one-warp
        CTAs driving the TMA unit (``snb_gather_rows_bulk``), 'stores' plain
        16-byte stores by `ctas` CTAs of 128 threads.  `src` may be the
        destination slice of the own buffer (rows produced in place): that
        copy is skipped.
        """
        import ctypes
        import torch
        from shennong_b200 import _lib
        if not src.is_contiguous():
            raise ValueError('contiguous rows expected')
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        src_p = ctypes.c_void_p(src.data_ptr())
        if how == 'ce':
            _lib.check(self._lib.snb_gather_rows_ce(
                src_p, src.numel(), self._ptrs, self._nptrs,
                int(offset_floats), stream))
        elif how == 'bulk':
            _lib.check(self._lib.snb_gather_rows_bulk(
                src_p, src.numel(), self._ptrs, self._nptrs,
                int(offset_floats), int(ctas), stream))
        elif how == 'stores':
            _lib.check(self._lib.snb_gather_rows(
                src_p, src.numel(), self._ptrs, self._nptrs,
                int(offset_floats), int(ctas) or 296, stream))
        else:
            raise ValueError(f'unknown collection method {how}')

    def arrive(self):
        """queued on the current stream: completes when every rank's pushes
        queued before its own arrive() have completed"""
        if self.size > 1:
            _dist().all_reduce(self._flag, group=self.group)

    def close(self):
        import ctypes
        import torch
        torch.cuda.synchronize()
        if self.size > 1:
            _dist().barrier(group=self.group)
        for p in self._opened:
            self._lib.snb_peer_buffer_close(ctypes.c_void_p(p))
        self._opened = []
        self.tensor = None
        if self._own:
            self._lib.snb_peer_buffer_destroy(ctypes.c_void_p(self._own))
            self._own = None


def _ceil4(n):
    return (int(n) + 3) // 4 * 4


class _LocalBuffer:
    """the result buffer of a rank when the collection is an NCCL all-gather
    (same interface as :class:`PeerGather`, nothing is mapped by the peers)"""
    fanout = 1

    def __init__(self, nfloats):
        import torch
        self.tensor = torch.empty(int(nfloats), dtype=torch.float32,
                                  device='cuda')

    def arrive(self):
        pass

    def close(self):
        self.tensor = None


class ChunkCollector:
    """Collection of a chunked, device-resident extraction inside the step

    Every rank runs the same :class:`FusedPipeline` over its own shard, cut in
    chunks of the same geometry on every rank (`chunk_offsets[k]` are the
    frame offsets of chunk k: equal shards, e.g. a corpus of equal-length
    utterances; ragged shards go through the chunk-wise NCCL gather of
    :mod:`shennong_b200.stream`).  The rows of chunk k are PRODUCED in this
    rank's block of the result buffer (:meth:`out_view`) and pushed to the
    same block of every peer's buffer while chunk k + 1 is computed
    (:class:`PeerGather`, own stream); :meth:`finish` closes the step: every
    rank then holds ``result(k)`` = [world, rows_k, D] for every chunk.

    `base_chunks` > 0 (pipelines with deltas, CMVN per utterance or none):
    for that many chunks only the BASE rows [rows_k, d] and the normalisation
    table of the chunk travel -- a third of the bytes for MFCC + delta +
    delta-delta -- and every receiver redoes the normalise + delta launch on
    them (``snb_cmvn_apply_deltas``, the same kernel on the same inputs: the
    rows are bit-identical to the sender's).  It trades NVLink time for
    HBM-bound recomputation: with 8 GPUs the all-gather of 39 columns (10.9 GB
    received per rank and step in BASELINE configs[2]) takes longer than the
    extraction itself.  The base-row chunks are the ones just before the last
    chunk: the links carry full rows from the first chunk on, the receivers'
    launches run at the end of the step under the push of the last chunk.
    """

    def __init__(self, pipe, chunk_offsets, how='ce', base_chunks=0, ctas=0,
                 fanout=1, group=None):
        import torch
        from shennong_b200 import engine
        self.pipe = pipe
        self.rank, self.size = world()
        self.how, self.ctas = how, int(ctas)
        self.offsets = [np.ascontiguousarray(o, dtype=np.int64)
                        for o in chunk_offsets]
        self.rows = [int(o[-1]) for o in self.offsets]
        self.nchunks = n = len(self.rows)
        D, d = pipe.out_dim, pipe.base_dim
        order = pipe.delta.order if pipe.delta is not None else 0
        nb = int(base_chunks)
        if nb and (order == 0 or pipe.cmvn == 'speaker'
                   or pipe.pitch is not None):
            raise ValueError(
                'base-rows collection needs a pipeline with deltas, without '
                'pitch columns and without CMVN by speaker')
        nb = min(nb, n) if self.size > 1 else 0
        first = max(n - 1 - nb, 0)
        self.base_ids = list(range(first, first + nb))
        # result buffer: chunk k = `size` blocks of stride[k] floats
        self.stride = [_ceil4(r * D) for r in self.rows]
        self.at = np.concatenate(
            ([0], np.cumsum([self.size * st for st in self.stride])))
        self.group = group
        nccl = how == 'nccl'
        self.peers = (_LocalBuffer(self.at[-1]) if nccl
                      else PeerGather(int(self.at[-1]), group, fanout))
        # NCCL gathers from a buffer of its own (no aliasing of input and
        # output); the peer pushes read the rank's block of the result
        self.own = None
        if nccl:
            self.own = [torch.empty(st, dtype=torch.float32, device='cuda')
                        for st in self.stride]
        self.staging = None
        self.bstride, self.nstride, self.bat = {}, {}, {}
        self.layouts, self.nutts = {}, {}
        if nb:
            at = 0
            for k in self.base_ids:
                self.nutts[k] = len(self.offsets[k]) - 1
                self.bstride[k] = _ceil4(self.rows[k] * d)
                self.nstride[k] = (_ceil4(self.nutts[k] * 2 * d)
                                   if pipe.cmvn else 0)
                self.bat[k] = at
                at += self.size * (self.bstride[k] + self.nstride[k])
                self.layouts[k] = engine.RowLayout(
                    frame_offsets=self.offsets[k])
            self.staging = (_LocalBuffer(at) if nccl
                            else PeerGather(at, group, fanout))
            if nccl:
                self.own_base = {k: torch.empty(
                    self.bstride[k] + self.nstride[k], dtype=torch.float32,
                    device='cuda') for k in self.base_ids}
        self.comm = torch.cuda.Stream()
        self._base_arrived = None

    # -- views ---------------------------------------------------------------
    def _block(self, k, r):
        a = int(self.at[k]) + r * self.stride[k]
        return self.peers.tensor[a:a + self.rows[k] * self.pipe.out_dim].view(
            self.rows[k], self.pipe.out_dim)

    def out_view(self, k):
        """[rows_k, D]: where this rank's pipeline writes chunk k"""
        if self.own is not None:
            return self.own[k][:self.rows[k] * self.pipe.out_dim].view(
                self.rows[k], self.pipe.out_dim)
        return self._block(k, self.rank)

    def result(self, k):
        """[world, rows_k, D] views of chunk k (valid after :meth:`finish`)"""
        return [self._block(k, r) for r in range(self.size)]

    def _base_block(self, k, r):
        a = self.bat[k] + r * self.bstride[k]
        d = self.pipe.base_dim
        return self.staging.tensor[a:a + self.rows[k] * d].view(
            self.rows[k], d)

    def _norm_block(self, k, r):
        if not self.nstride[k]:
            return None
        d = self.pipe.base_dim
        a = self.bat[k] + self.size * self.bstride[k] + r * self.nstride[k]
        return self.staging.tensor[a:a + self.nutts[k] * 2 * d].view(
            self.nutts[k], 2, d)

    def base_view(self, k):
        """[rows_k, d] buffer for the base rows of chunk k (``base_buf`` of
        ``run_device``) when the chunk travels as base rows, else None"""
        if k not in self.bat:
            return None
        if self.own is not None:
            d = self.pipe.base_dim
            return self.own_base[k][:self.rows[k] * d].view(self.rows[k], d)
        return self._base_block(k, self.rank)

    def norm_view(self, k):
        """``norm_out`` of ``run_device`` for chunk k (or None)"""
        if k not in self.bat or not self.nstride[k]:
            return None
        if self.own is not None:
            d = self.pipe.base_dim
            a = self.bstride[k]
            return self.own_base[k][a:a + self.nutts[k] * 2 * d].view(
                self.nutts[k], 2, d)
        return self._norm_block(k, self.rank)

    # -- the step --------------------------------------------------------------
    def collect(self, k):
        """Queues the push of chunk k behind the work queued so far on the
        current stream (call right after ``run_device`` of the chunk)"""
        import torch
        if self.size == 1 and self.peers.fanout == 1:
            return
        cur = torch.cuda.current_stream()
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(done)
            if self.own is not None:
                self._collect_nccl(k)
            elif k in self.bat:
                a = self.bat[k] + self.rank * self.bstride[k]
                src = self.staging.tensor[a:a + self.bstride[k]]
                self.staging.push(src, a, ctas=self.ctas, how=self.how)
                if self.nstride[k]:
                    a = (self.bat[k] + self.size * self.bstride[k]
                         + self.rank * self.nstride[k])
                    src = self.staging.tensor[a:a + self.nstride[k]]
                    self.staging.push(src, a, ctas=self.ctas, how=self.how)
                if k == self.base_ids[-1]:
                    self.staging.arrive()
            else:
                a = int(self.at[k]) + self.rank * self.stride[k]
                src = self.peers.tensor[a:a + self.stride[k]]
                self.peers.push(src, a, ctas=self.ctas, how=self.how)
            if self.base_ids and k == self.base_ids[-1]:
                self._base_arrived = torch.cuda.Event()
                self._base_arrived.record(self.comm)

    def _collect_nccl(self, k):
        """all-gathers of chunk k (the collective completes with the data of
        every rank in place: no separate arrival)"""
        dist = _dist()
        if k in self.bat:
            a, n = self.bat[k], self.bstride[k]
            dist.all_gather_into_tensor(
                self.staging.tensor[a:a + self.size * n],
                self.own_base[k][:n], group=self.group)
            if self.nstride[k]:
                a, m = a + self.size * n, self.nstride[k]
                dist.all_gather_into_tensor(
                    self.staging.tensor[a:a + self.size * m],
                    self.own_base[k][n:n + m], group=self.group)
            # this rank's final rows of the chunk stay local
            self._block(k, self.rank).copy_(self.out_view(k))
        else:
            a, n = int(self.at[k]), self.stride[k]
            dist.all_gather_into_tensor(
                self.peers.tensor[a:a + self.size * n], self.own[k],
                group=self.group)

    def finish(self):
        """Queues the end of the step on the current stream: the receivers'
        normalise + delta launches on the base rows of the other ranks, then
        the arrival of every rank's pushes"""
        import torch
        from shennong_b200 import engine
        cur = torch.cuda.current_stream()
        if self.size == 1:
            cur.wait_stream(self.comm)
            return
        if self.base_ids:
            cur.wait_event(self._base_arrived)
            pipe = self.pipe
            order, window = pipe.delta.order, pipe.delta.window
            for k in self.base_ids:
                for r in range(self.size):
                    if r == self.rank:
                        continue
                    engine.deltas(
                        self._base_block(k, r), self.layouts[k], order,
                        window, norm=self._norm_block(k, r),
                        out=self._block(k, r)[:, :pipe.feat_dim])
            # (a peer may overwrite the staging only after these launches:
            # the closing all-reduce waits for them)
            self.comm.wait_stream(cur)
        with torch.cuda.stream(self.comm):
            self.peers.arrive()
        cur.wait_stream(self.comm)

    def close(self):
        self.peers.close()
        if self.staging is not None:
            self.staging.close()


def allreduce_stats(stats, group=None):
    """Sums float64 CMVN statistics [G, 2, d+1] over the ranks (only needed
    when a speaker is split across GPUs)"""
    dist = _dist()
    if world()[1] > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def extract_sharded(pipe, signals, speakers=None, gather=True):
    """Runs a :class:`FusedPipeline` on this rank's shard of `signals`

    Every rank passes the same full list (or each rank may pass ``None`` for
    the utterances it does not own: only the owned ones are read).  Returns
    (features of ALL utterances [total_frames, D] when `gather`, else the
    local block; per-utterance frame counts in the ORIGINAL order; the
    utterance indices in the order of the returned rows).
    """
    from shennong_b200 import _lib, engine
    rank, size = world()
    fo = pipe.processor._frame_opts()
    lengths = np.array([len(s) if s is not None else 0 for s in signals])
    if size > 1:
        # lengths of non-owned utterances may be unknown locally: agree on them
        import torch
        dist = _dist()
        t = torch.from_numpy(lengths.astype(np.int64))
        if dist.get_backend() == 'nccl':
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lengths = t.cpu().numpy()
    frames = engine.num_frames_array(fo, lengths).astype(np.int64)
    groups = speakers if pipe.cmvn == 'speaker' else None
    shards = shard_utterances(frames, size, groups)
    mine = shards[rank]
    packed = engine.PackedAudio([signals[i] for i in mine])
    local, offs, _, _ = pipe.run_device(
        packed, speakers=[speakers[i] for i in mine] if speakers else None)
    order = np.concatenate(shards) if size > 1 else mine
    if not gather or size == 1:
        return local, frames, (order if gather else mine)
    full, _ = gather_rows(local)
    return full, frames, order
