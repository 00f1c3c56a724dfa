"""Streamed extraction (shennong_b200/stream.py) on the GPU: whatever the
chunking, the blocks of speakers and the PCM source, the rows equal those of
one resident batch (`FusedPipeline.run_device`), which the other GPU tests
hold against the oracle."""

import numpy as np
import pytest
import scipy.io.wavfile
import torch

import oracle
from conftest import scale_close, synth_utterance
from shennong_b200 import Audio, Utterances, engine, pipeline, stream
from shennong_b200.fused import FusedPipeline
from shennong_b200.postprocessor import DeltaPostProcessor, VadPostProcessor
from shennong_b200.processor import (
    EnergyProcessor, FilterbankProcessor, KaldiPitchPostProcessor,
    KaldiPitchProcessor, MfccProcessor, PlpProcessor)

pytestmark = pytest.mark.gpu

LENGTHS = [48000, 16000, 80000, 22713, 160000, 9000, 31999, 64000, 5000,
           22157, 399, 40000, 16000, 72000]


def packed_corpus(lengths):
    sigs = [synth_utterance(i, n) for i, n in enumerate(lengths)]
    padded = [(n + 7) // 8 * 8 for n in lengths]
    starts = np.concatenate(([0], np.cumsum(padded)))[:-1].astype(np.int64)
    host = torch.zeros(int(sum(padded)) + 64, dtype=torch.int16,
                       pin_memory=True)
    view = host.numpy()
    for s, sig in zip(starts, sigs):
        view[s:s + len(sig)] = sig
    return sigs, host, starts, np.asarray(lengths, dtype=np.int64)


def pitch_pair():
    return (KaldiPitchProcessor(),
            KaldiPitchPostProcessor(delta_pitch_noise_stddev=0))


@pytest.mark.parametrize('chunk_utts', [1, 3, 512])
@pytest.mark.parametrize('case', ['mfcc_cmvn_delta', 'fbank', 'plp_pitch'])
def test_run_host_equals_resident_batch(case, chunk_utts):
    sigs, host, starts, lengths = packed_corpus(LENGTHS)
    pipe = {
        'mfcc_cmvn_delta': lambda: FusedPipeline(
            MfccProcessor(dither=0), delta=DeltaPostProcessor(),
            cmvn='utterance'),
        'fbank': lambda: FusedPipeline(FilterbankProcessor(dither=0)),
        'plp_pitch': lambda: FusedPipeline(
            PlpProcessor(dither=0), pitch=pitch_pair())}[case]()
    out, foffs = pipe.run_host(host, starts, lengths, chunk_utts=chunk_utts)
    ref, offs, stats, _ = pipe.run_device(engine.PackedAudio(sigs))
    assert np.array_equal(foffs, offs)
    ref = ref.cpu().numpy()
    valid = pipe.valid_rows
    for u in range(len(sigs)):
        n = int(valid[u]) if valid is not None else int(offs[u + 1] - offs[u])
        a = int(offs[u])
        assert np.array_equal(out.numpy()[a:a + n], ref[a:a + n]), (case, u)
    if case == 'mfcc_cmvn_delta':
        assert np.array_equal(pipe.host_stats, stats.cpu().numpy())


@pytest.mark.parametrize('block_rows', [300, 900, 10**9])
def test_speaker_cmvn_blocks(block_rows):
    """CMVN by speaker with VAD weights, delta and pitch: blocks of whole
    speakers (two-pass) equal the resident batch"""
    sigs, host, starts, lengths = packed_corpus(LENGTHS)
    speakers = ['s%02d' % (i // 3) for i in range(len(sigs))]
    pipe = FusedPipeline(
        FilterbankProcessor(dither=0), delta=DeltaPostProcessor(),
        cmvn='speaker', vad=VadPostProcessor(),
        energy=EnergyProcessor(dither=0), pitch=pitch_pair())
    runner = stream.StreamRunner(
        pipe, chunk_utts=2, block_bytes=4 * (23 + 72) * block_rows)
    out, plan, stats = runner.run(
        stream.PackedSource(host, starts, lengths), speakers=speakers)
    if block_rows < 1000:
        assert len(plan.blocks) > 1
    ref, offs, rstats, _ = pipe.run_device(
        engine.PackedAudio(sigs), speakers=speakers)
    ref = ref.cpu().numpy()
    assert out.shape == ref.shape == (offs[-1], 72)
    for u in range(len(sigs)):
        a, n = int(offs[u]), int(plan.valid[u])
        assert np.array_equal(out.numpy()[a:a + n], ref[a:a + n]), u
    assert np.array_equal(stats, rstats.cpu().numpy())
    assert runner.group_names == sorted(set(speakers))


@pytest.fixture(scope='module')
def wav_corpus(tmp_path_factory):
    root = tmp_path_factory.mktemp('stream_wavs')
    entries = []
    for i, n in enumerate(LENGTHS):
        if n < 400:
            continue
        path = root / f'w{i}.wav'
        scipy.io.wavfile.write(path, 16000, synth_utterance(i, n))
        entries.append((f'utt{i:02d}', str(path), 'spk%d' % (i % 4)))
    # two segments of the longest file
    entries.append(('seg_a', str(root / 'w4.wav'), 'spk0', 1.0, 3.5))
    entries.append(('seg_b', str(root / 'w4.wav'), 'spk1', 0.0, 0.73))
    return entries


def test_process_all_streams_wav_files(wav_corpus, monkeypatch):
    monkeypatch.setenv('SNB_STREAM_CHUNK_UTTS', '4')
    # (one format for all: whole files as the segment [0, duration])
    utts = Utterances([
        (e[0], e[1]) + (tuple(e[3:]) or (0.0, Audio.scan(e[1]).duration))
        for e in wav_corpus])
    proc = MfccProcessor(dither=0)
    feats = proc.process_all(utts, njobs=3)
    assert set(feats.keys()) == {e[0] for e in wav_corpus}
    for utt in utts:
        ref = proc.process(utt.load_audio())
        got = feats[utt.name]
        assert got.shape == ref.shape
        assert np.array_equal(got.data, ref.data), utt.name
        assert np.array_equal(got.times, ref.times)
        assert got.properties == ref.properties
        assert got.is_valid()
    # warps by utterance
    warps = {u.name: 0.9 + 0.02 * i for i, u in enumerate(utts)}
    feats = proc.process_all(utts, vtln_warp=warps)
    for utt in list(utts)[::4]:
        ref = proc.process(utt.load_audio(), vtln_warp=warps[utt.name])
        assert np.array_equal(feats[utt.name].data, ref.data)
        assert feats[utt.name].properties['mfcc']['vtln_warp'] == \
            warps[utt.name]
    with pytest.raises(ValueError, match='sample rates'):
        MfccProcessor(sample_rate=8000).process_all(utts)


def test_extract_features_streamed_small_chunks(wav_corpus, monkeypatch):
    """full default pipeline (CMVN by speaker with VAD, delta, pitch) with
    chunks of 3 utterances and blocks of ~one speaker"""
    config = pipeline.get_default_config(
        'mfcc', with_pitch='kaldi', with_cmvn=True, with_delta=True)
    config['mfcc']['dither'] = 0
    config['pitch']['postprocessing']['delta_pitch_noise_stddev'] = 0
    # (one format for all: whole files as the segment [0, duration])
    utts = Utterances([
        tuple(e[:3]) + (tuple(e[3:]) or (0.0, Audio.scan(e[1]).duration))
        for e in wav_corpus])
    ref = pipeline.extract_features(config, utts)
    monkeypatch.setenv('SNB_STREAM_CHUNK_UTTS', '3')
    monkeypatch.setenv('SNB_STREAM_BLOCK_BYTES', str(4 * (13 + 42) * 1200))
    got = pipeline.extract_features(config, utts, njobs=4)
    assert list(got.keys()) == list(ref.keys())
    for name in ref:
        assert got[name].shape == ref[name].shape
        assert got[name].shape[1] == 42
        assert np.array_equal(got[name].data, ref[name].data), name
        assert got[name].properties['speaker'] == ref[name].properties[
            'speaker']
        assert np.array_equal(got[name].properties['cmvn']['stats'],
                              ref[name].properties['cmvn']['stats'])
    # against the oracle: column block 0 = CMVN by speaker (VAD-weighted) of
    # the MFCC, then deltas
    by_spk = {}
    for utt in utts:
        by_spk.setdefault(utt.speaker, []).append(utt)
    for spk, members in by_spk.items():
        mf = [oracle.features('mfcc', u.load_audio().data, dither=0)
              for u in members]
        st = sum(oracle.cmvn_accumulate(
            m, oracle.vad(oracle.features(
                'energy', u.load_audio().data, dither=0).astype(
                    np.float32).reshape(-1, 1)).reshape(-1).astype(np.float32))
            for m, u in zip(mf, members))
        for m, u in zip(mf, members):
            want = oracle.deltas(oracle.cmvn_apply(m, st))
            scale_close(got[u.name].data[:, :39], want, 2e-4)


def test_cmvn_needs_one_frame(tmp_path):
    """an utterance too short for a frame: the reference's CMVN raises on the
    empty statistics (cmvn.py:254-257)"""
    paths = []
    for i, n in enumerate([16000, 300]):
        paths.append(str(tmp_path / f'{i}.wav'))
        scipy.io.wavfile.write(paths[-1], 16000, synth_utterance(i, n))
    config = pipeline.get_default_config('mfcc', with_cmvn=True)
    config['cmvn']['by_speaker'] = False
    utts = Utterances([('a', paths[0]), ('b', paths[1])])
    with pytest.raises(ValueError, match='insufficient accumulation'):
        pipeline.extract_features(config, utts)
