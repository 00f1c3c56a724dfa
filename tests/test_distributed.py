"""Multi-process host logic on CPU: sharding and ragged gather over gloo,
world_size 2 (the N>1 GPU path uses the same code over NCCL)."""

import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from shennong_b200.distributed import shard_utterances


def test_shards_are_contiguous_balanced_partitions():
    rng = np.random.default_rng(0)
    costs = rng.integers(100, 2000, 1000)
    for world in (1, 2, 3, 8):
        shards = shard_utterances(costs, world)
        assert len(shards) == world
        assert np.array_equal(np.concatenate(shards), np.arange(1000))
        loads = np.array([costs[s].sum() for s in shards])
        assert loads.max() - loads.min() <= 2 * costs.max()
    # equal costs: equal shards
    shards = shard_utterances(np.full(10000, 998), 8)
    assert [len(s) for s in shards] == [1250] * 8
    # more ranks than utterances
    shards = shard_utterances([5, 5], 4)
    assert sorted(np.concatenate(shards).tolist()) == [0, 1]
    assert shard_utterances([], 2)[0].size == 0


def test_shards_keep_speakers_together():
    rng = np.random.default_rng(1)
    speakers = [f'spk{rng.integers(0, 37)}' for _ in range(500)]
    costs = rng.integers(100, 1000, 500)
    for world in (2, 4, 8):
        shards = shard_utterances(costs, world, groups=speakers)
        assert sorted(np.concatenate(shards).tolist()) == list(range(500))
        owner = {}
        for rank, shard in enumerate(shards):
            for i in shard:
                assert owner.setdefault(speakers[i], rank) == rank
        loads = np.array([costs[s].sum() for s in shards])
        assert loads.max() < 2.0 * costs.sum() / world


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ['SNB_ROOT'])
    from shennong_b200.distributed import (
        allreduce_stats, gather_rows, shard_utterances, world)
    dist.init_process_group('gloo')
    rank, size = world()
    assert size == 2
    # every rank derives the same partition from the same costs
    costs = np.arange(1, 11) * 100
    shards = shard_utterances(costs, size)
    mine = shards[rank]
    # a "feature" block whose rows identify (utterance, frame)
    rows = [np.stack([np.full(c // 100, u), np.arange(c // 100)], 1)
            for u, c in enumerate(costs)]
    local = torch.from_numpy(
        np.concatenate([rows[u] for u in mine]).astype(np.float32))
    full, counts = gather_rows(local)
    expect = np.concatenate([rows[u] for s in shards for u in s])
    assert np.array_equal(full.numpy(), expect), rank
    assert counts.sum() == expect.shape[0] and len(counts) == 2
    stats = torch.full((3, 2, 5), float(rank + 1), dtype=torch.float64)
    assert torch.all(allreduce_stats(stats) == 3.0)
    # empty shard on one rank
    local = torch.zeros((0 if rank else 4, 3))
    full, counts = gather_rows(local)
    assert full.shape == (4, 3) and counts.tolist() == [4, 0]
    dist.destroy_process_group()
    print('worker-%d-ok' % rank, flush=True)
''')


def test_gather_over_gloo_world_size_2(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    env = dict(os.environ, SNB_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
         '--nproc-per-node=2', '--master-addr', '127.0.0.1',
         '--master-port', str(port), str(script)],
        env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count('-ok') == 2 and 'worker-0' in out.stdout
