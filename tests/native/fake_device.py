"""TEST INFRASTRUCTURE: numpy stand-ins for the few CUDA-tensor operations the tests and bench.py use.
In the fake CUDA runtime 'device memory' is host memory, so a numpy array is a device buffer."""
import numpy as np


class _Dev:
    """Stand-in for a CUDA uint8 tensor: in the fake runtime 'device memory' is host memory, so a numpy array
    with the handful of tensor methods the tests use is enough."""
    def __init__(self, a):
        self.a = a

    def cuda(self):
        return self

    def cpu(self):
        return self

    def numpy(self):
        return self.a

    def data_ptr(self):
        return self.a.ctypes.data

    def zero_(self):
        self.a[:] = 0
        return self

    def __getitem__(self, k):
        return _Dev(self.a[k])


class _Torch:
    uint8 = np.uint8

    @staticmethod
    def from_numpy(a):
        return _Dev(np.ascontiguousarray(a).copy())

    @staticmethod
    def empty(n, dtype=None, device=None):
        return _Dev(np.empty(n, dtype=np.uint8))
