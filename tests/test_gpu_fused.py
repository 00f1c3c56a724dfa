"""Fused device pipeline (features -> cmvn -> delta [+ pitch]) vs the
per-processor API and the oracle; chunked multi-stream host path."""

import numpy as np
import pytest

import oracle
from conftest import cmvn_close, scale_close, synth_utterance
from shennong_b200 import Audio, engine
from shennong_b200.fused import FusedPipeline
from shennong_b200.postprocessor import DeltaPostProcessor, VadPostProcessor
from shennong_b200.processor import (
    EnergyProcessor, KaldiPitchPostProcessor, KaldiPitchProcessor,
    MfccProcessor)

pytestmark = pytest.mark.gpu

LENGTHS = [160000, 16000, 48000, 400, 22713, 100, 80000, 31999, 160000]


@pytest.fixture(scope='module')
def signals():
    return [synth_utterance(i, n) for i, n in enumerate(LENGTHS)]


def oracle_pipeline(sig, cmvn=True, order=2):
    base = oracle.features('mfcc', sig)
    if base.shape[0] == 0:
        return np.zeros((0, 13 * (order + 1)), np.float32)
    if cmvn:
        base = oracle.cmvn_apply(base, oracle.cmvn_accumulate(base))
    return oracle.deltas(base, order, 2) if order else base


def test_device_pipeline_matches_oracle(signals):
    pipe = FusedPipeline(MfccProcessor(dither=0),
                         delta=DeltaPostProcessor(order=2, window=2),
                         cmvn='utterance')
    packed = engine.PackedAudio(signals)
    out, offs, stats, _ = pipe.run_device(packed)
    out = engine.to_host(out)
    assert out.shape == (offs[-1], 39)
    for i, sig in enumerate(signals):
        if offs[i + 1] - offs[i] < 3:
            continue    # variance of 1-2 frames is floored: x * 1e10 noise
        cmvn_close(out[offs[i]:offs[i + 1]], oracle_pipeline(sig),
                   oracle.features('mfcc', sig))
    st = engine.to_host(stats)
    assert st.shape == (len(signals), 2, 14)
    assert np.array_equal(st[:, 0, -1], np.diff(offs).astype(np.float64))


def test_speaker_cmvn_and_vad(signals):
    speakers = ['a', 'b', 'a', 'c', 'b', 'c', 'a', 'b', 'c']
    vad = VadPostProcessor()
    pipe = FusedPipeline(
        MfccProcessor(dither=0), delta=DeltaPostProcessor(order=1, window=2),
        cmvn='speaker', vad=vad, energy=EnergyProcessor(dither=0))
    out, offs, stats, group = pipe.run_device(
        engine.PackedAudio(signals), speakers=speakers)
    out, stats = engine.to_host(out), engine.to_host(stats)
    for g, spk in enumerate(['a', 'b', 'c']):
        ref_stats = np.zeros((2, 14))
        for sig, s in zip(signals, speakers):
            if s != spk:
                continue
            base = oracle.features('mfcc', sig)
            if base.shape[0] == 0:
                continue
            w = oracle.vad(oracle.features('energy', sig).astype(np.float32))
            ref_stats = oracle.cmvn_accumulate(
                base, w[:, 0].astype(np.float32), ref_stats)
        # (float32 features differ by ~1e-6 between GPU and oracle)
        assert np.allclose(stats[g], ref_stats, rtol=1e-5, atol=1e-2)
        for i, (sig, s) in enumerate(zip(signals, speakers)):
            if s != spk or offs[i + 1] - offs[i] < 3:
                continue
            base = oracle.features('mfcc', sig)
            ref = oracle.deltas(oracle.cmvn_apply(base, ref_stats), 1, 2)
            count = ref_stats[0, -1]
            sigma = np.sqrt(ref_stats[1, :-1] / count
                            - (ref_stats[0, :-1] / count) ** 2)
            cmvn_close(out[offs[i]:offs[i + 1]], ref, base, sigma=sigma)


def test_pitch_columns(signals):
    sigs = [signals[0], signals[2], signals[4]]
    pitch = (KaldiPitchProcessor(),
             KaldiPitchPostProcessor(delta_pitch_noise_stddev=0))
    pipe = FusedPipeline(MfccProcessor(dither=0),
                         delta=DeltaPostProcessor(), cmvn='utterance',
                         pitch=pitch)
    out, offs, _, _ = pipe.run_device(engine.PackedAudio(sigs))
    out = engine.to_host(out)
    assert out.shape[1] == 42
    for i, sig in enumerate(sigs):
        block = out[offs[i]:offs[i + 1]]
        cmvn_close(block[:, :39], oracle_pipeline(sig),
                   oracle.features('mfcc', sig))
        raw = pitch[0].process(Audio(sig, 16000))
        ref = pitch[1].process(raw).data
        assert np.allclose(block[:, 39:], ref, atol=1e-5)


@pytest.mark.parametrize('chunk', [1, 2, 4, 100])
def test_run_host_chunked_streams(signals, chunk):
    import torch
    pipe = FusedPipeline(MfccProcessor(dither=0),
                         delta=DeltaPostProcessor(order=2, window=2),
                         cmvn='utterance')
    packed = engine.PackedAudio(signals)
    ref, offs, _, _ = pipe.run_device(packed)
    ref = engine.to_host(ref)
    out, foffs = pipe.run_host(
        packed.host, packed.starts, packed.lengths, chunk_utts=chunk)
    torch.cuda.synchronize()
    assert np.array_equal(foffs, offs)
    assert np.array_equal(out.numpy(), ref)
    # repeated calls reuse plans and give identical results
    out2, _ = pipe.run_host(
        packed.host, packed.starts, packed.lengths, chunk_utts=chunk)
    assert np.array_equal(out2.numpy(), ref)


def test_stream_ordered_batches_recycle_safely(signals):
    """Batches are created on the current stream and their device blob is
    recycled through the library's pool: destroying a batch right after
    queueing the kernels that read it (here on a side stream, many times,
    with changing shapes) must never corrupt a launch still in flight"""
    import torch
    proc = MfccProcessor(dither=0)
    plan = engine.feature_plan(
        proc._frame_opts(), proc._mel_opts(), proc._feat_opts())
    subsets = [signals[:k] for k in (9, 3, 5, 1, 7, 2)]
    packs = [engine.PackedAudio(s) for s in subsets]
    refs = []
    for p in packs:
        refs.append(engine.to_host(
            engine.compute_features(plan, engine.Batch(plan, p))))
    side = torch.cuda.Stream()
    outs = []
    with torch.cuda.stream(side):
        for rep in range(40):
            p = packs[rep % len(packs)]
            batch = engine.Batch(plan, p)
            outs.append((rep % len(packs),
                         engine.compute_features(plan, batch)))
            del batch            # blob goes back to the pool, kernel queued
    side.synchronize()
    for k, out in outs:
        assert np.array_equal(engine.to_host(out), refs[k])
