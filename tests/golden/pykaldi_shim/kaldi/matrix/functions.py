import numpy as np
import scipy.fft


def vec_vec(a, b):
    """VecVec<float>: cblas_sdot (float32 accumulation)"""
    return float(np.dot(a.numpy().astype(np.float32), b.numpy().astype(np.float32)))


def real_fft(vec, forward):
    """RealFft(v, true): in place, layout [re0, re(N/2), re1, im1, re2, im2, ...] in float32"""
    assert forward
    x = vec.numpy()
    n = x.shape[0]
    spec = scipy.fft.rfft(x.astype(np.float32))
    assert spec.dtype == np.complex64
    x[0] = spec[0].real
    x[1] = spec[n // 2].real
    x[2::2] = spec[1:n // 2].real
    x[3::2] = spec[1:n // 2].imag
