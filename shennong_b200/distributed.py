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
    at the same offset of ALL of them with one libsnb kernel
    (``snb_gather_rows``) on the current stream -- the all-gather of the
    north star as posted writes, issued by the producer as soon as a chunk is
    done, with no rendezvous per chunk.  :meth:`arrive` is the one
    synchronisation of a step: a 4-byte all-reduce queued behind the pushes.
    """

    def __init__(self, nfloats, group=None):
        import ctypes
        import torch
        from shennong_b200 import _lib
        dist = _dist()
        self.rank, self.size = world()
        self.group = group
        self.nfloats = int(nfloats)
        L = _lib.lib()
        self._lib = L
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(L.snb_peer_buffer_create(
            self.nfloats * 4, ctypes.byref(ptr), handle))
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
            ptrs.append(peer.value)
        self._ptrs = (ctypes.c_void_p * self.size)(*ptrs)
        self.tensor = torch.as_tensor(
            _DeviceBuffer(self._own, self.nfloats), device='cuda')
        self._flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        if self.size > 1:
            dist.barrier(group=group)

    def push(self, src, offset_floats, ctas=296):
        """`src` (contiguous float32 device tensor, a multiple of 4 elements)
        -> floats [offset, offset + src.numel()) of every rank's buffer"""
        import ctypes
        import torch
        from shennong_b200 import _lib
        if not src.is_contiguous():
            raise ValueError('contiguous rows expected')
        _lib.check(self._lib.snb_gather_rows(
            ctypes.c_void_p(src.data_ptr()), src.numel(), self._ptrs,
            self.size, int(offset_floats), int(ctas),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

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
