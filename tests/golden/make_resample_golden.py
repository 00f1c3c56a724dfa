#!/usr/bin/env python
"""Golden vectors for the sample-rate converter

    python tests/golden/make_resample_golden.py      (needs torchaudio; run in the build container)

Kaldi's LinearResample (resample.cc) is not available offline; its published
algorithm is restated by ``torchaudio.functional.resample`` with
``resampling_method='sinc_interp_hann'``, ``lowpass_filter_width=6`` and
``rolloff=0.99`` -- the parameters of kaldi::ResampleWaveform (torchaudio's
implementation was written as a port of it).  The vectors are computed in
float64 on the first 8 000 samples of the reference's test/data/test.wav (the
``pcm`` array of kaldi_compliance.npz) and stored as float32.
"""
import os

import numpy as np
import torch
import torchaudio.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
PAIRS = [(16000, 8000), (16000, 44100), (16000, 11025), (44100, 16000),
         (8000, 16000), (16000, 4000)]


def main():
    pcm = np.load(os.path.join(HERE, 'kaldi_compliance.npz'))['pcm'][:8000]
    out = {'pcm': pcm.astype(np.int16)}
    x = torch.from_numpy(pcm.astype(np.float64))[None]
    for rate_in, rate_out in PAIRS:
        y = F.resample(x, rate_in, rate_out, lowpass_filter_width=6,
                       rolloff=0.99, resampling_method='sinc_interp_hann')
        out[f'{rate_in}_{rate_out}'] = y[0].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, 'resample_sinc_hann.npz'), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == '__main__':
    main()
