"""Write / read times of the on-disk formats on a collection of the size the
reference quotes its own numbers on (shennong/features_collection.py:12-26:
MFCC of the Buckeye corpus, 38 h, 254 files, 883.7 MB as pickle).

The collection is what ``pipeline.extract_features`` returns: one Features per
utterance whose data is a VIEW into the single (pinned) result matrix of the
streamed extraction, timestamps built lazily.  CPU only.

    python tools/bench_serializers.py [--hours 38] [--files 254] [--csv]
"""

import argparse
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

REFERENCE = {          # features_collection.py:17-26 (size, write s, read s)
    'pickle': ('883.7 MB', 7, 5), 'numpy': ('869.1 MB', 150, 22),
    'matlab': ('721.1 MB', 59, 11), 'kaldi': ('1.3 GB', 6, 7),
    'csv': ('4.8 GB', 182, 191)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--hours', type=float, default=38.0)
    ap.add_argument('--files', type=int, default=254)
    ap.add_argument('--csv', action='store_true')
    ap.add_argument('--dir', default=None)
    args = ap.parse_args()
    from shennong_b200 import Features, FeaturesCollection
    from shennong_b200.processor import MfccProcessor
    proc = MfccProcessor()
    frames_total = int(args.hours * 3600 * 100)
    per = frames_total // args.files
    rng = np.random.default_rng(0)
    big = rng.standard_normal((per * args.files, 13)).astype(np.float32)
    props = proc.get_properties(vtln_warp=1.0)
    coll = FeaturesCollection(
        ('utt%04d' % i, Features(
            big[i * per:(i + 1) * per], proc.times(per),
            dict(props, audio={'file': '/corpus/s%04d.wav' % i}),
            validate=False)) for i in range(args.files))
    root = args.dir or tempfile.mkdtemp(prefix='snb_ser_')
    formats = [('pickle', 'feats.pkl'), ('numpy', 'feats.npz'),
               ('matlab', 'feats.mat'), ('kaldi', 'feats.ark')]
    if args.csv:
        formats.append(('csv', 'feats_csv'))
    print(f'{args.files} files, {per * args.files} frames x 13 float32 '
          f'({big.nbytes / 1e6:.1f} MB of data), directory {root}')
    print(f'{"format":8s} {"size":>10s} {"write s":>9s} {"read s":>8s}   '
          f'reference (size, write, read)')
    for name, fname in formats:
        path = os.path.join(root, fname)
        t0 = time.perf_counter()
        coll.save(path, serializer=name)
        tw = time.perf_counter() - t0
        if os.path.isdir(path):
            size = sum(os.path.getsize(os.path.join(path, f))
                       for f in os.listdir(path))
        else:
            size = os.path.getsize(path)
            extra = path.replace('.ark', '.times.ark')
            if name == 'kaldi':
                size += sum(os.path.getsize(os.path.join(root, f))
                            for f in os.listdir(root)
                            if f.startswith('feats.') and f != fname
                            and ('ark' in f or 'scp' in f or 'json' in f
                                 or 'properties' in f))
        t0 = time.perf_counter()
        back = FeaturesCollection.load(path, serializer=name)
        tr = time.perf_counter() - t0
        assert len(back) == len(coll)
        first = next(iter(coll))
        assert np.array_equal(back[first].data, coll[first].data)
        ref = REFERENCE[name]
        print(f'{name:8s} {size / 1e6:8.1f}MB {tw:9.2f} {tr:8.2f}   '
              f'{ref[0]}, {ref[1]} s, {ref[2]} s', flush=True)
    if args.dir is None:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
    main()
