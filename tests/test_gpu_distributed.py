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
    offset = sum(1000 + 8 * r for r in range(rank)) * dim
    for round_, how in enumerate(('ce', 'bulk', 'stores', 'ce')):
        base = (rank + 1) * 1000.0 + 10000.0 * round_
        block = base + torch.arange(
            rows * dim, dtype=torch.float32, device='cuda').view(rows, dim) / 7
        if round_ == 3:
            # rows produced in place in the own buffer: only the peers are written
            mine = peers.tensor[offset:offset + rows * dim].view(rows, dim)
            mine.copy_(block)
            block = mine
        peers.push(block, offset, ctas=8, how=how)
        peers.arrive()
        torch.cuda.synchronize()
        got = peers.tensor.clone()
        at = 0
        for r in range(size):
            n = (1000 + 8 * r) * dim
            want = (r + 1) * 1000.0 + 10000.0 * round_ + torch.arange(
                n, dtype=torch.float32, device='cuda') / 7
            assert torch.equal(got[at:at + n], want), (rank, r, how)
            at += n
        dist.barrier()
    peers.close()
    # ---- chunked device-resident step with the collection inside --------------
    from shennong_b200.distributed import ChunkCollector
    n_utt, n_samp, n_chunk = 12, 32000, 3
    mine = [synth_utterance(100 + rank * n_utt + i, n_samp) for i in range(n_utt)]
    pipe = FusedPipeline(MfccProcessor(dither=0), delta=DeltaPostProcessor(),
                         cmvn='utterance')
    plans = pipe._plans()
    per = n_utt // n_chunk
    packs = [engine.PackedAudio(mine[k * per:(k + 1) * per]) for k in range(n_chunk)]
    batches = [pipe.make_batches(plans, p) for p in packs]
    want_local, _, _, _ = pipe.run_device(engine.PackedAudio(mine))
    want = [torch.empty_like(want_local) for _ in range(size)]
    dist.all_gather(want, want_local.contiguous())
    for how, nb in (('ce', 0), ('ce', 2), ('bulk', 3), ('stores', 1),
                    ('nccl', 0), ('nccl', 2)):
        coll = ChunkCollector(
            pipe, [b['feat'].frame_offsets for b in batches], how=how,
            base_chunks=nb, ctas=8)
        for rep in range(2):         # twice: buffers are reused across steps
            for k in range(n_chunk):
                pipe.run_device(packs[k], out=coll.out_view(k), plans=plans,
                                base_buf=coll.base_view(k), batches=batches[k],
                                norm_out=coll.norm_view(k))
                coll.collect(k)
            coll.finish()
            torch.cuda.synchronize()
            for r in range(size):
                got = torch.cat([coll.result(k)[r] for k in range(n_chunk)])
                assert torch.equal(got, want[r]), (how, nb, rep, rank, r)
            if rep == 0:             # poison: the second step must rewrite all
                dist.barrier()
                for k in range(n_chunk):
                    for r in range(size):
                        coll.result(k)[r].fill_(float('nan'))
                torch.cuda.synchronize()
                dist.barrier()
        coll.close()
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
