#!/usr/bin/env python
"""Where the time of the reference-facing batch calls goes (host side)

    python tools/profile_api.py [--utts 2000]

Writes 10 s WAV files to a temporary directory, warms the path up, then runs
``MfccProcessor.process_all`` and ``pipeline.extract_features`` under cProfile
and prints the functions with the largest cumulative time, plus wall-clock
times of the stages of shennong_b200.stream (scan, read into pinned staging).
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

import numpy as np
import scipy.io.wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--utts', type=int, default=2000)
    ap.add_argument('--top', type=int, default=22)
    args = ap.parse_args()
    import torch
    from shennong_b200 import Utterances, pipeline, stream
    from shennong_b200.processor import MfccProcessor
    root = tempfile.mkdtemp(prefix='snb_prof_')
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(160000) * 1000).astype(np.int16)
    for u in range(args.utts):
        scipy.io.wavfile.write(os.path.join(root, 'u%05d.wav' % u), 16000, x)
    utts = Utterances([('u%05d' % u, os.path.join(root, 'u%05d.wav' % u),
                        'spk%03d' % (u // 100)) for u in range(args.utts)])
    cores = len(os.sched_getaffinity(0))
    proc = MfccProcessor()
    config = pipeline.get_default_config('mfcc', with_cmvn=True, with_delta=True)
    calls = {
        'process_all': lambda: proc.process_all(utts, njobs=cores),
        'extract_features': lambda: pipeline.extract_features(
            config, utts, njobs=cores)}
    for name, call in calls.items():
        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        prof = cProfile.Profile()
        prof.enable()
        call()
        prof.disable()
        out = io.StringIO()
        pstats.Stats(prof, stream=out).sort_stats('cumulative').print_stats(args.top)
        print(f'==== {name}: {dt * 1e3:.1f} ms for {args.utts} utterances '
              f'({cores} cores)')
        print('\n'.join(out.getvalue().splitlines()[4:]))
    ulist = [utts[k] for k in utts.by_name().keys()]
    t0 = time.perf_counter()
    items, lengths, _ = stream.audio_items(ulist, sample_rate=16000)
    t1 = time.perf_counter()
    print(f'audio_items (layouts cached): {(t1 - t0) * 1e3:.1f} ms')
    for workers in (1, 4, cores):
        src = stream.AudioSource(items, lengths, workers=workers)
        staging = torch.empty(src.span(0, 512), dtype=torch.int16, pin_memory=True)
        for rep in range(2):
            t0 = time.perf_counter()
            for b in range(0, args.utts, 512):
                src.window(b, min(b + 512, args.utts), staging)
            dt = time.perf_counter() - t0
        print(f'read into pinned staging, {workers} workers: {dt * 1e3:.1f} ms '
              f'({lengths.sum() * 2 / dt / 1e9:.1f} GB/s)')


if __name__ == '__main__':
    main()
