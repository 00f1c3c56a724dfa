import math

import numpy as np
import torchaudio.compliance.kaldi as ta_kaldi


class MelBanksOptions:
    def __init__(self):
        self.num_bins = 23
        self.low_freq = 20.0
        self.high_freq = 0.0
        self.vtln_low = 100.0
        self.vtln_high = -500.0


class MelBanks:
    """mel triangles from torchaudio's Kaldi-compliance implementation (independent of this repo)"""
    def __init__(self, mel_opts, frame_opts, vtln_warp):
        bins, centers = ta_kaldi.get_mel_banks(
            int(mel_opts.num_bins), int(frame_opts.padded_window_size()), float(frame_opts.samp_freq),
            float(mel_opts.low_freq), float(mel_opts.high_freq), float(mel_opts.vtln_low),
            float(mel_opts.vtln_high), float(vtln_warp))
        self._bins = bins.numpy().astype(np.float32)        # [num_bins, N/2]
        self._centers = centers.numpy().astype(np.float32)

    def num_bins(self):
        return self._bins.shape[0]

    def compute(self, power_spectrum, mel_energies_out):
        p = power_spectrum.numpy()[:self._bins.shape[1]].astype(np.float32)
        mel_energies_out.numpy()[:] = self._bins @ p


def compute_lifter_coeffs(q, coeffs):
    """ComputeLifterCoeffs (mel-computations.cc)"""
    c = coeffs.numpy()
    for i in range(c.shape[0]):
        c[i] = 1.0 + 0.5 * q * math.sin(math.pi * i / q)


def get_equal_loudness_vector(mel_banks):
    """GetEqualLoudnessVector (mel-computations.cc), BaseFloat = float"""
    from ..matrix import Vector
    f32 = np.float32
    f0 = mel_banks._centers
    out = np.zeros(f0.shape[0], dtype=np.float32)
    for i in range(f0.shape[0]):
        fsq = f32(f0[i] * f0[i])
        fsub = f32(fsq / f32(fsq + f32(1.6e5)))
        out[i] = f32(f32(fsub * fsub) * f32(f32(fsq + f32(1.44e6)) / f32(fsq + f32(9.61e6))))
    return Vector._view(out)


def compute_lpc(autocorr, lpc_out):
    """ComputeLpc + Durbin (mel-computations.cc), all float32; returns -Log(1/E)"""
    f32 = np.float32
    ac = autocorr.numpy()
    lp = lpc_out.numpy()
    n = ac.shape[0] - 1
    tmp = np.zeros(n, dtype=np.float32)
    e = f32(ac[0])
    for i in range(n):
        ki = f32(ac[i + 1])
        for j in range(i):
            ki = f32(ki + f32(lp[j] * ac[i - j]))
        ki = f32(ki / e)
        c = f32(f32(1) - f32(ki * ki))
        if c < f32(1.0e-5):
            c = f32(1.0e-5)
        e = f32(e * c)
        tmp[i] = -ki
        for j in range(i):
            tmp[j] = f32(lp[j] - f32(ki * lp[i - j - 1]))
        lp[:i + 1] = tmp[:i + 1]
    return float(f32(-math.log(1.0 / float(e))))   # -Log(1.0 / ans): double arithmetic, float result
