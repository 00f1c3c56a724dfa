"""Secondary measurements on one GPU for the other BASELINE.json configs
(device-resident PCM, CUDA events around each of 5 reps after 3 warm-ups, median;
host-side batch creation is inside the timed call; not the bench.py line).

    python tools/bench_configs.py [--utts N]
"""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--utts', type=int, default=2000)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    import torch
    import bench
    from shennong_b200 import engine
    from shennong_b200.fused import FusedPipeline
    from shennong_b200.postprocessor import (
        DeltaPostProcessor, VadPostProcessor)
    from shennong_b200.processor import (
        EnergyProcessor, FilterbankProcessor, KaldiPitchPostProcessor,
        KaldiPitchProcessor, MfccProcessor, PlpProcessor,
        SpectrogramProcessor)

    n = args.utts
    pcm = torch.cat([bench.synth_pcm_device(n, 0, torch),
                     torch.zeros(64, dtype=torch.int16, device='cuda')])
    starts = np.arange(n, dtype=np.int64) * bench.UTT_SAMPLES
    lengths = np.full(n, bench.UTT_SAMPLES, dtype=np.int64)
    packed = engine.PackedAudio.from_packed(None, starts, lengths, dev=pcm)
    speakers = [f'spk{i // 100}' for i in range(n)]
    pitch = (KaldiPitchProcessor(), KaldiPitchPostProcessor())

    cases = {
        'cfg1_fbank40': FusedPipeline(FilterbankProcessor(num_bins=40)),
        'cfg1_fbank40_dither0': FusedPipeline(
            FilterbankProcessor(num_bins=40, dither=0)),
        'mfcc_only': FusedPipeline(MfccProcessor()),
        'spectrogram': FusedPipeline(SpectrogramProcessor()),
        'plp_only': FusedPipeline(PlpProcessor()),
        'rasta_plp': FusedPipeline(PlpProcessor(rasta=True)),
        'cfg3_mfcc_delta_cmvn': FusedPipeline(
            MfccProcessor(), delta=DeltaPostProcessor(), cmvn='utterance'),
        'pitch_only': FusedPipeline(MfccProcessor(), pitch=pitch),
        'cfg4_plp_pitch': FusedPipeline(PlpProcessor(), pitch=pitch),
        'cfg5_fbank_pitch_delta_cmvn_spk_vad': FusedPipeline(
            FilterbankProcessor(), delta=DeltaPostProcessor(),
            cmvn='speaker', vad=VadPostProcessor(), energy=EnergyProcessor(),
            pitch=pitch),
        # generic kernel (any FFT size): the same PCM read as 8 kHz audio, N = 256
        'mfcc_8k_generic_path': FusedPipeline(MfccProcessor(sample_rate=8000)),
    }
    # the VTLN trainer's warp grid as ONE fused batch of
    # warps x utterances virtual utterances over the same PCM (21 warps,
    # 0.85 .. 1.25, pipeline.extract_features_warp_sweep)
    grid = np.round(np.arange(0.85, 1.2501, 0.02), 2).astype(np.float32)
    nsweep = max(1, n // len(grid))
    sweep_packed = engine.PackedAudio.from_packed(
        None, np.tile(starts[:nsweep], len(grid)),
        np.tile(lengths[:nsweep], len(grid)), dev=pcm)
    sweep_warps = np.repeat(grid, nsweep)
    cases['vtln_sweep_21warps_mfcc_delta'] = FusedPipeline(
        MfccProcessor(), delta=DeltaPostProcessor())
    # ragged batch (load balance): the same buffer read as utterances of 2..10 s
    ragged_lengths = np.random.default_rng(7).integers(
        32000, bench.UTT_SAMPLES + 1, size=n).astype(np.int64)
    ragged_packed = engine.PackedAudio.from_packed(
        None, starts, ragged_lengths, dev=pcm)
    cases['cfg3_ragged_2_10s_mfcc_delta_cmvn'] = FusedPipeline(
        MfccProcessor(), delta=DeltaPostProcessor(), cmvn='utterance')
    results = {}
    for name, pipe in cases.items():
        if args.only and args.only not in name:
            continue
        if pipe is None:
            continue
        plans = pipe._plans()

        def run():
            if name.startswith('vtln_sweep'):
                return pipe.run_device(sweep_packed, warps=sweep_warps,
                                       plans=plans)
            if 'ragged' in name:
                return pipe.run_device(ragged_packed, speakers=speakers,
                                       plans=plans)
            return pipe.run_device(packed, speakers=speakers, plans=plans)
        for _ in range(3):
            out, offs, _, _ = run()
        torch.cuda.synchronize()
        reps = 5
        times = []
        for _ in range(reps):
            e0, e1 = (torch.cuda.Event(enable_timing=True),
                      torch.cuda.Event(enable_timing=True))
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        frames = int(offs[-1])
        results[name] = {'ms': ms, 'ms_min': min(times), 'ms_max': max(times),
                         'frames': frames, 'dim': int(out.shape[1]),
                         'frames_per_s': frames / (ms * 1e-3)}
        print(f'{name:40s} {ms:9.3f} ms (min {min(times):.3f} max '
              f'{max(times):.3f})  {frames / (ms * 1e-3):.3e} frames/s '
              f'D={out.shape[1]}', flush=True)
    print(json.dumps(results))


if __name__ == '__main__':
    main()
