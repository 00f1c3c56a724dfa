import math

import numpy as np

from ..matrix import Matrix


def compute_power_spectrum(vec):
    """ComputePowerSpectrum: RealFft layout -> power in the first N/2+1 entries"""
    x = vec.numpy()
    n = x.shape[0]
    half = n // 2
    first = np.float32(x[0] * x[0])
    last = np.float32(x[1] * x[1])
    re = x[2::2].copy()
    im = x[3::2].copy()
    x[1:half] = re * re + im * im
    x[0] = first
    x[half] = last


def init_idft_bases(n_bases, dimension):
    """InitIdftBases (feature-functions.cc), BaseFloat = float"""
    f32 = np.float32
    angle = f32(math.pi / f32(dimension - 1))
    scale = f32(1.0 / (2.0 * f32(dimension - 1)))
    out = np.zeros((n_bases, dimension), dtype=np.float32)
    for i in range(n_bases):
        out[i, 0] = f32(1.0 * scale)
        i_fl = f32(i)
        for j in range(1, dimension - 1):
            j_fl = f32(j)
            out[i, j] = f32(2.0 * float(scale) * math.cos(float(f32(f32(angle * i_fl) * j_fl))))
        out[i, dimension - 1] = f32(float(scale) * math.cos(float(f32(f32(angle * i_fl) * f32(dimension - 1)))))
    return Matrix._from(out)
