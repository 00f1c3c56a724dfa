"""Host logic of the streamed extraction (shennong_b200/stream.py): chunk and
block planning, the PCM sources, WAV layout parsing -- no GPU needed -- and
the chunk-wise all-gather of the collection step over gloo, world_size 2."""

import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import scipy.io.wavfile

from conftest import ROOT
from shennong_b200 import Audio, stream
from shennong_b200.fused import FusedPipeline
from shennong_b200.postprocessor import DeltaPostProcessor
from shennong_b200.processor import (
    KaldiPitchPostProcessor, KaldiPitchProcessor, MfccProcessor)


def test_plan_chunks_cover_the_corpus_in_order():
    rng = np.random.default_rng(0)
    lengths = rng.integers(300, 200000, 1000)
    pipe = FusedPipeline(MfccProcessor(), delta=DeltaPostProcessor(),
                         cmvn='utterance')
    plan = stream.StreamPlan(pipe, lengths, chunk_utts=64,
                             chunk_samples=3_000_000)
    assert plan.chunks[0][0] == 0 and plan.chunks[-1][1] == 1000
    padded = (lengths + 7) // 8 * 8
    for (b, e), (b2, _) in zip(plan.chunks, plan.chunks[1:] + [(1000, 0)]):
        assert e == b2 and 0 < e - b <= 64
        assert e - b == 1 or padded[b:e].sum() <= 3_000_000
    # rows: Kaldi's frame count, snip_edges
    frames = np.where(lengths < 400, 0, 1 + (lengths - 400) // 160)
    assert np.array_equal(plan.frames, frames)
    assert plan.total == frames.sum()
    assert plan.blocks == [(i, i + 1) for i in range(len(plan.chunks))]
    assert plan.max_chunk_rows == max(
        frames[b:e].sum() for b, e in plan.chunks)


def test_plan_blocks_hold_whole_speakers():
    rng = np.random.default_rng(1)
    groups = np.sort(rng.integers(0, 40, 2000))
    lengths = rng.integers(16000, 160000, 2000)
    pipe = FusedPipeline(MfccProcessor(), delta=DeltaPostProcessor(),
                         cmvn='speaker')
    row_bytes = 4 * (13 + 39)
    plan = stream.StreamPlan(pipe, lengths, groups, chunk_utts=100,
                             block_bytes=row_bytes * 150000)
    assert len(plan.blocks) > 3
    covered = 0
    for first, last in plan.blocks:
        ub, ue = plan.chunks[first][0], plan.chunks[last - 1][1]
        assert ub == covered
        covered = ue
        # no speaker is split over two blocks
        if ue < 2000:
            assert groups[ue - 1] != groups[ue]
        rows = plan.foffs[ue] - plan.foffs[ub]
        one_speaker = groups[ub] == groups[ue - 1]
        assert rows <= 150000 or one_speaker
    assert covered == 2000


def test_plan_pitch_trims_to_the_shorter():
    """pitch has one frame more than the features for some lengths: the rows
    to keep are the shorter of the two (Features.concatenate(tolerance=2))"""
    pipe = FusedPipeline(MfccProcessor(), pitch=(
        KaldiPitchProcessor(), KaldiPitchPostProcessor()))
    lengths = np.arange(22000, 22400)
    plan = stream.StreamPlan(pipe, lengths)
    assert np.all(plan.valid <= plan.frames)
    assert np.all(plan.frames - plan.valid <= 2)
    import oracle
    for n in (22000, 22157, 22399):
        assert plan.frames[n - 22000] == oracle.num_frames(n)


def test_wav_layout_and_audio_source(tmp_path):
    rng = np.random.default_rng(2)
    sigs = [rng.integers(-2000, 2000, n).astype(np.int16)
            for n in (16000, 401, 22713, 8)]
    paths = []
    for i, sig in enumerate(sigs):
        paths.append(str(tmp_path / f'{i}.wav'))
        scipy.io.wavfile.write(paths[-1], 16000, sig)
    fpath = str(tmp_path / 'f.wav')
    scipy.io.wavfile.write(fpath, 16000, (sigs[0] / 2**15).astype(np.float32))
    assert Audio.wav_layout(fpath) is None           # float: not a raw read
    stereo = str(tmp_path / 's.wav')
    scipy.io.wavfile.write(stereo, 16000, np.stack([sigs[0], sigs[0]], 1))
    assert Audio.wav_layout(stereo) is None
    items, lengths = [], []
    for path, sig in zip(paths, sigs):
        offset, n, rate = Audio.wav_layout(path)
        assert n == len(sig) and rate == 16000
        items.append((path, offset, 0, n))
        lengths.append(n)
    # the native header walk (snb_wav_scan_batch) against the Python one,
    # incl. files it must refuse, a missing file and an extra chunk before
    # `data`
    listed = str(tmp_path / 'listed.wav')
    raw = open(paths[1], 'rb').read()
    at = raw.index(b'data')
    extra = b'LIST' + (6).to_bytes(4, 'little') + b'abcdef'
    body = raw[:at] + extra + raw[at:]
    body = body[:4] + (len(body) - 8).to_bytes(4, 'little') + body[8:]
    open(listed, 'wb').write(body)
    probe = paths + [fpath, stereo, str(tmp_path / 'missing.wav'), listed,
                     paths[0]]
    assert stream.wav_layouts(probe, nthreads=3) == [
        Audio.wav_layout(p) for p in probe]
    assert stream.wav_layouts(probe)[-2][1] == len(sigs[1])
    # a truncated file is reported like the reference's loader does
    short = str(tmp_path / 'short.wav')
    open(short, 'wb').write(open(paths[0], 'rb').read()[:-100])
    off, n, _ = stream.wav_layouts([short])[0]
    with pytest.raises(ValueError, match='cannot read file'):
        src = stream.AudioSource([(short, off, 0, n + 50)], [n + 50])
        import torch
        src.window(0, 1, torch.zeros(src.span(0, 1), dtype=torch.int16))
    # a segment of a file, an in-memory int16 array and a float Audio
    offset = Audio.wav_layout(paths[0])[0]
    items += [(paths[0], offset, 1000, 5000), sigs[2],
              Audio((sigs[0] / 2**15).astype(np.float32), 16000)]
    lengths += [5000, len(sigs[2]), 16000]
    source = stream.AudioSource(items, lengths, workers=3)
    import torch
    staging = torch.zeros(source.span(0, len(items)), dtype=torch.int16)
    window, rel = source.window(0, len(items), staging)
    got = window.numpy()
    expect = sigs + [sigs[0][1000:6000], sigs[2], sigs[0]]
    assert np.all(rel % 8 == 0)
    for r, sig in zip(rel, expect):
        assert np.array_equal(got[r:r + len(sig)], sig)


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ['SNB_ROOT'])
    from shennong_b200 import stream
    from shennong_b200.distributed import shard_utterances, world
    from shennong_b200.fused import FusedPipeline
    from shennong_b200.processor import MfccProcessor
    dist.init_process_group('gloo')
    rank, size = world()
    pipe = FusedPipeline(MfccProcessor())
    rng = np.random.default_rng(3)
    lengths = rng.integers(300, 60000, 57)
    frames = np.where(lengths < 400, 0, 1 + (lengths - 400) // 160)
    shards = shard_utterances(frames, size)
    # rank 1 gets smaller chunks: it has more pieces than rank 0
    plans = [stream.StreamPlan(pipe, lengths[s], chunk_utts=7 - 3 * r)
             for r, s in enumerate(shards)]
    total = sum(p.total for p in plans)
    out = torch.zeros((total, 13))
    coll = stream.GatherCollector(out, None, None, plans, rank, 13)
    # the "features" of row i of rank r: r * 1e6 + i in every column
    plan = plans[rank]
    for c in range(len(plan.chunks)):
        a, b = plan.chunk_rows(c)
        rows = (rank * 1e6 + torch.arange(a, b, dtype=torch.float32))
        coll.put(rows[:, None].repeat(1, 13), a, None)
    coll.finish()
    expect = np.concatenate([r * 1e6 + np.arange(p.total)
                             for r, p in enumerate(plans)])
    assert np.array_equal(out.numpy()[:, 0], expect.astype(np.float32)), rank
    assert np.array_equal(out.numpy()[:, 12], expect.astype(np.float32))
    assert coll.pieces == max(len(p.chunks) for p in plans)
    dist.destroy_process_group()
    print('worker-%d-ok' % rank, flush=True)
''')


def test_chunkwise_gather_over_gloo_world_size_2(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
         '--nproc-per-node=2', '--master-addr', '127.0.0.1',
         '--master-port', str(port), str(script)],
        env=dict(os.environ, SNB_ROOT=ROOT), capture_output=True, text=True,
        timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count('-ok') == 2
