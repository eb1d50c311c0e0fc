"""CPU: the Python mirror's argument checks (no GPU needed: borrowed arrays are never dereferenced here)."""
import numpy as np
import pytest

from cupoch_b200.utility import DeviceArray, Matrix3fVector, Vector3fVector


class FakeCuda:
    def __init__(self, shape, typestr, strides=None):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (0x1000, False), "version": 2,
                                         "strides": strides}


def test_borrowed_cuda_arrays_must_be_float32():
    # kernels read raw float32 memory: a float64 / float16 array must be rejected, not reinterpreted
    assert Vector3fVector(FakeCuda((5, 3), "<f4")).dtype == np.float32
    with pytest.raises(ValueError):
        Vector3fVector(FakeCuda((5, 3), "<f8"))
    with pytest.raises(ValueError):
        Vector3fVector(FakeCuda((5, 3), "<f2"))
    with pytest.raises(ValueError):
        Matrix3fVector(FakeCuda((5, 3, 3), "<f8"))
    with pytest.raises(ValueError):
        Vector3fVector(FakeCuda((5, 4), "<f4"))                      # wrong shape
    with pytest.raises(ValueError):
        Vector3fVector(FakeCuda((5, 3), "<f4", strides=(16, 4)))     # not C-contiguous (float4 rows)
    assert DeviceArray.wrap(FakeCuda((7, 2), "<i4"), np.int32).dtype == np.int32
    with pytest.raises(ValueError):
        DeviceArray.wrap(FakeCuda((7, 2), "<i8"), np.int32)


def test_torch_cuda_tensors_are_converted_not_reinterpreted():
    torch = pytest.importorskip("torch")

    class T:  # stands in for a CUDA tensor (no GPU here): only the attributes borrow() looks at
        is_cuda = True

        def __init__(self, dtype, shape=(4, 3)):
            self.dtype, self.shape, self.converted = dtype, shape, False

        def data_ptr(self):
            return 0x2000

        def is_contiguous(self):
            return True

        def to(self, dt):
            t = T(dt, self.shape)
            t.converted = True
            return t

    d = Vector3fVector(T(torch.float64))
    assert d.dtype == np.float32 and d._base.converted and d._base.dtype == torch.float32
    d = Vector3fVector(T(torch.float32))
    assert d.dtype == np.float32 and not d._base.converted
