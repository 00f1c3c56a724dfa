"""N>1 GPU path: speaker-aligned sharding + NCCL all-gather of the features,
one process per GPU under torchrun (skipped with fewer than 2 GPUs)."""

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
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ['SNB_ROOT'])
    sys.path.insert(0, os.path.join(os.environ['SNB_ROOT'], 'tests'))
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl')
    from conftest import synth_utterance
    from shennong_b200 import engine
    from shennong_b200.distributed import extract_sharded, world
    from shennong_b200.fused import FusedPipeline
    from shennong_b200.postprocessor import DeltaPostProcessor
    from shennong_b200.processor import MfccProcessor
    rank, size = world()
    lengths = [48000, 16000, 80000, 22713, 160000, 9000, 31999, 64000, 5000]
    signals = [synth_utterance(i, n) for i, n in enumerate(lengths)]
    speakers = ['s%d' % (i % 3) for i in range(len(lengths))]
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
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
         '--nproc-per-node=2', '--master-addr', '127.0.0.1',
         '--master-port', str(port), str(script)],
        env=dict(os.environ, SNB_ROOT=ROOT), capture_output=True, text=True,
        timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count('-ok') == 2
