"""N>1 GPU path, one process per GPU under torchrun (skipped with fewer than
2 GPUs): speaker-aligned sharding, the collection by NCCL all-gather
(device-resident and chunk-wise in the streamed host API) and by direct
stores into peer memory (PeerGather), against single-GPU results."""

import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import scipy.io.wavfile
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ['SNB_ROOT'])
    sys.path.insert(0, os.path.join(os.environ['SNB_ROOT'], 'tests'))
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl')
    from conftest import synth_utterance
    from shennong_b200 import Utterances, engine, pipeline
    from shennong_b200.distributed import (
        PeerGather, extract_sharded, world)
    from shennong_b200.fused import FusedPipeline
    from shennong_b200.postprocessor import DeltaPostProcessor
    from shennong_b200.processor import MfccProcessor
    rank, size = world()
    lengths = [48000, 16000, 80000, 22713, 160000, 9000, 31999, 64000, 5000]
    signals = [synth_utterance(i, n) for i, n in enumerate(lengths)]
    speakers = ['s%d' % (i % 3) for i in range(len(lengths))]
    # ---- device-resident sharding + all-gather -------------------------------
    for cmvn in ('utterance', 'speaker'):
        pipe = FusedPipeline(MfccProcessor(dither=0), delta=DeltaPostProcessor(),
                             cmvn=cmvn)
        full, frames, order = extract_sharded(pipe, signals, speakers=speakers)
        full = full.cpu().numpy()
        # single-process result of the same utterances, in the gathered order
        ref, offs, _, _ = pipe.run_device(
            engine.PackedAudio([signals[i] for i in order]),
            speakers=[speakers[i] for i in order])
        ref = ref.cpu().numpy()
        assert full.shape == ref.shape == (frames.sum(), 39), (full.shape, ref.shape)
        assert np.array_equal(full, ref), cmvn
    # ---- peer-memory gather: every rank pushes a block to all ranks ----------
    rows, dim = 1000 + 8 * rank, 12
    total = sum(1000 + 8 * r for r in range(size)) * dim
    peers = PeerGather(total)
    block = (rank + 1) * 1000.0 + torch.arange(
        rows * dim, dtype=torch.float32, device='cuda').view(rows, dim) / 7
    offset = sum(1000 + 8 * r for r in range(rank)) * dim
    peers.push(block, offset, ctas=8)
    peers.arrive()
    torch.cuda.synchronize()
    got = peers.tensor.clone()
    at = 0
    for r in range(size):
        n = (1000 + 8 * r) * dim
        want = (r + 1) * 1000.0 + torch.arange(
            n, dtype=torch.float32, device='cuda') / 7
        assert torch.equal(got[at:at + n], want), (rank, r)
        at += n
    peers.close()
    # ---- the host API, sharded: every rank gets the whole collection ---------
    root = os.environ['SNB_WAVS']
    if rank == 0:
        for i, sig in enumerate(signals):
            scipy.io.wavfile.write(os.path.join(root, 'w%d.wav' % i), 16000, sig)
    dist.barrier()
    utts = Utterances([('utt%d' % i, os.path.join(root, 'w%d.wav' % i),
                        speakers[i]) for i in range(len(signals))])
    config = pipeline.get_default_config(
        'mfcc', with_pitch='kaldi', with_cmvn=True, with_delta=True)
    config['mfcc']['dither'] = 0
    config['pitch']['postprocessing']['delta_pitch_noise_stddev'] = 0
    os.environ['SNB_STREAM_CHUNK_UTTS'] = '2'
    feats = pipeline.extract_features(config, utts)
    assert len(feats) == len(signals)
    # reference: this rank alone on the whole corpus (resident batch)
    manager = pipeline.PipelineManager(config, utts)
    first = next(iter(utts))
    pipe = FusedPipeline(
        MfccProcessor(dither=0), delta=DeltaPostProcessor(), cmvn='speaker',
        vad=manager.get_vad_processor(),
        energy=manager.get_energy_processor(first),
        pitch=(manager.get_pitch_processor(first),
               manager.get_pitch_post_processor()))
    ref, offs, _, _ = pipe.run_device(
        engine.PackedAudio(signals), speakers=speakers)
    ref = ref.cpu().numpy()
    for i in range(len(signals)):
        got = feats['utt%d' % i]
        want = ref[offs[i]:offs[i] + got.shape[0]]
        assert got.shape[1] == 42 and np.array_equal(got.data, want), i
        assert got.properties['speaker'] == speakers[i]
    own = pipeline.extract_features(config, utts, gather=False)
    assert 0 < len(own) < len(signals)
    for name, f in own.items():
        assert np.array_equal(f.data, feats[name].data)
    mf = MfccProcessor(dither=0).process_all(utts)
    for i in range(len(signals)):
        one = MfccProcessor(dither=0).process(utts['utt%d' % i].load_audio())
        assert np.array_equal(mf['utt%d' % i].data, one.data), i
    dist.barrier()
    dist.destroy_process_group()
    print('rank-%d-ok' % rank, flush=True)
''')


def test_sharded_extraction_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    wavs = tmp_path / 'wavs'
    wavs.mkdir()
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
         '--nproc-per-node=2', '--master-addr', '127.0.0.1',
         '--master-port', str(port), str(script)],
        env=dict(os.environ, SNB_ROOT=ROOT, SNB_WAVS=str(wavs)),
        capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count('-ok') == 2
