"""Vector / SubVector / Matrix over float32 numpy views (aliasing like Kaldi's)."""
import numpy as np

from . import common, functions  # noqa: F401


class Vector:
    def __init__(self, arg=None):
        if arg is None:
            self._a = np.zeros(0, dtype=np.float32)
        elif isinstance(arg, (int, np.integer)):
            self._a = np.zeros(int(arg), dtype=np.float32)
        elif isinstance(arg, Vector):
            self._a = arg._a.copy()
        else:
            self._a = np.array(arg, dtype=np.float32).reshape(-1)

    @classmethod
    def _view(cls, array):
        v = cls.__new__(cls)
        v._a = array
        return v

    @property
    def dim(self):
        return int(self._a.shape[0])

    def numpy(self):
        return self._a

    def resize_(self, n, resize_type=None):
        self._a = np.zeros(int(n), dtype=np.float32)

    def __len__(self):
        return self.dim

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return Vector._view(self._a[idx])
        return float(self._a[idx])

    def __setitem__(self, idx, value):
        if isinstance(value, Vector):
            value = value._a
        self._a[idx] = value

    def sum(self):
        # VectorBase<float>::Sum accumulates in float
        acc = np.float32(0)
        for v in self._a:
            acc = np.float32(acc + v)
        return float(acc)

    def add_(self, value):
        self._a += np.float32(value)

    def scale_(self, value):
        self._a *= np.float32(value)

    def set_zero_(self):
        self._a[:] = 0

    def mul_elements_(self, other):
        self._a *= other._a if isinstance(other, Vector) else np.asarray(other, dtype=np.float32)

    def apply_pow_(self, power):
        self._a[:] = np.power(self._a, np.float32(power), dtype=np.float32)

    def add_mat_vec_(self, alpha, mat, trans, vec, beta):
        out = mat._a.astype(np.float32) @ vec._a.astype(np.float32)
        self._a[:] = np.float32(beta) * self._a + np.float32(alpha) * out.astype(np.float32)


class SubVector(Vector):
    def __init__(self, data):
        self._a = np.asarray(data, dtype=np.float32).reshape(-1)


class Matrix:
    def __init__(self, rows=0, cols=0):
        self._a = np.zeros((int(rows), int(cols)), dtype=np.float32)

    @classmethod
    def _from(cls, array):
        m = cls.__new__(cls)
        m._a = np.asarray(array, dtype=np.float32)
        return m

    def numpy(self):
        return self._a

    def __getitem__(self, idx):
        out = self._a[idx]
        if out.ndim == 1:
            return Vector._view(out)
        return Matrix._from(out)


SubMatrix = Matrix
