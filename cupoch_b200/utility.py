"""Device containers mirroring cupoch's utility vectors (device_vector_wrapper.h:34-58):
Vector3fVector(numpy) uploads; .cpu() downloads.  Also borrows torch CUDA tensors /
anything with __cuda_array_interface__ without copying (the reference does this through
DLPack, utility/dl_converter.h:32-40)."""
import ctypes as C

import numpy as np

from . import _lib


class DeviceArray:
    """A typed, shaped view of device memory.  Owns it unless it borrows from `base`."""

    def __init__(self, shape, dtype, ptr=None, base=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._base = base
        if ptr is None:
            _lib.require_gpu()
            ptr = _lib.lib().cphb_malloc(max(self.nbytes, 16))
            if not ptr:
                _lib.check(-2)
            self._owned = True
        else:
            self._owned = False
        self.ptr = int(ptr)

    # -- construction ------------------------------------------------------
    @classmethod
    def from_numpy(cls, a, dtype=np.float32):
        a = np.ascontiguousarray(a, dtype=dtype)
        d = cls(a.shape, dtype)
        if a.nbytes:
            _lib.check(_lib.lib().cphb_memcpy_h2d(d.ptr, a.ctypes.data, a.nbytes, None))
            _lib.check(_lib.lib().cphb_stream_synchronize(None))
        return d

    @classmethod
    def borrow(cls, obj, dtype=None):
        """Zero-copy view of a torch CUDA tensor / CuPy array (must be contiguous).  With `dtype` the element type is
        ENFORCED: the kernels read raw float32 / int32 memory, so a float64 tensor must never be handed over as it is
        -- torch tensors of another dtype are converted (a copy, kept alive by the view), other CUDA arrays are
        rejected."""
        want = None if dtype is None else np.dtype(dtype)
        if hasattr(obj, "data_ptr") and hasattr(obj, "is_cuda"):
            if not obj.is_cuda:
                raise ValueError("need a CUDA tensor")
            names = {"torch.float32": np.float32, "torch.int32": np.int32, "torch.float64": np.float64,
                     "torch.int64": np.int64, "torch.uint8": np.uint8, "torch.float16": np.float16}
            have = names.get(str(obj.dtype))
            if want is not None and (have is None or np.dtype(have) != want):
                import torch
                tdt = {"float32": torch.float32, "int32": torch.int32, "float64": torch.float64, "int64": torch.int64,
                       "uint8": torch.uint8}.get(want.name)
                if tdt is None:
                    raise ValueError("cannot convert a %s tensor to %s" % (obj.dtype, want))
                obj, have = obj.to(tdt), want.type
            if have is None:
                raise ValueError("unsupported tensor dtype %s" % obj.dtype)
            if not obj.is_contiguous():
                obj = obj.contiguous()
            return cls(tuple(obj.shape), have, ptr=obj.data_ptr(), base=obj)
        cai = obj.__cuda_array_interface__
        have = np.dtype(cai["typestr"])
        if want is not None and have != want:
            raise ValueError("expected a %s CUDA array, got %s" % (want, have))
        if cai.get("strides") is not None:
            st, expect = tuple(cai["strides"]), []
            acc = have.itemsize
            for d in reversed(tuple(cai["shape"])):
                expect.insert(0, acc)
                acc *= d
            if st != tuple(expect):
                raise ValueError("need a C-contiguous CUDA array")
        return cls(cai["shape"], have, ptr=cai["data"][0], base=obj)

    @classmethod
    def wrap(cls, obj, dtype=np.float32):
        if obj is None:
            return obj
        if isinstance(obj, DeviceArray):
            if dtype is not None and obj.dtype != np.dtype(dtype):
                raise ValueError("expected a %s device array, got %s" % (np.dtype(dtype), obj.dtype))
            return obj
        if hasattr(obj, "is_cuda") and obj.is_cuda:
            return cls.borrow(obj, dtype)
        if hasattr(obj, "__cuda_array_interface__"):
            return cls.borrow(obj, dtype)
        if hasattr(obj, "numpy") and not isinstance(obj, np.ndarray):
            obj = obj.numpy()
        return cls.from_numpy(np.asarray(obj), dtype)

    @property
    def __cuda_array_interface__(self):
        """zero-copy export (torch.as_tensor(arr, device="cuda"), cupy.asarray(arr)); the reference exports through
        DLPack (utility/dl_converter.h:32-40)"""
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 2, "strides": None}

    # -- access --------------------------------------------------------------
    def cpu(self, count=None):
        """download (optionally only the first `count` rows)"""
        shape = self.shape if count is None else (int(count),) + self.shape[1:]
        out = np.empty(shape, self.dtype)
        if out.nbytes:
            L = _lib.lib()
            if out.nbytes >= (1 << 16):
                stage = _PinnedStage.get(out.nbytes)
                _lib.check(L.cphb_memcpy_d2h(stage, self.ptr, out.nbytes, None))
                _lib.check(L.cphb_stream_synchronize(None))
                C.memmove(out.ctypes.data, stage, out.nbytes)
            else:
                _lib.check(L.cphb_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, None))
                _lib.check(L.cphb_stream_synchronize(None))
        return out

    def __len__(self):
        return self.shape[0] if self.shape else 0

    def __del__(self):
        if getattr(self, "_owned", False) and self.ptr:
            try:
                _lib.lib().cphb_free(self.ptr)
            except Exception:
                pass
            self.ptr = 0


class _PinnedStage:
    """One reusable pinned host buffer for D2H reads (pageable copies run at a few GB/s)."""
    ptr, cap = 0, 0

    @classmethod
    def get(cls, nbytes):
        if nbytes > cls.cap:
            if cls.ptr:
                _lib.lib().cphb_free_host(cls.ptr)
            cls.cap = max(nbytes, 1 << 20)
            cls.ptr = _lib.lib().cphb_malloc_host(cls.cap)
            if not cls.ptr:
                cls.cap = 0
                _lib.check(-2)
        return cls.ptr


class _DevicePool:
    """Tiny size-keyed free list so per-call result buffers do not hit cudaMalloc/cudaFree."""
    free = {}

    @classmethod
    def take(cls, shape, dtype):
        key = (tuple(shape), np.dtype(dtype).str)
        lst = cls.free.get(key)
        return lst.pop() if lst else DeviceArray(shape, dtype)

    @classmethod
    def give(cls, arr):
        if arr is not None and arr._owned and arr.ptr:
            key = (arr.shape, arr.dtype.str)
            lst = cls.free.setdefault(key, [])
            if len(lst) < 4:
                lst.append(arr)


def Vector3fVector(a=None):
    if a is None:
        return None
    d = DeviceArray.wrap(a, np.float32)
    if len(d.shape) != 2 or d.shape[1] != 3:
        raise ValueError("expected (n,3) float32")
    return d


def Matrix3fVector(a=None):
    if a is None:
        return None
    d = DeviceArray.wrap(a, np.float32)
    if d.shape[1:] not in ((3, 3), (9,)):
        raise ValueError("expected (n,3,3) float32")
    return d


def as_f16(T):
    T = np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(4, 4))
    return (C.c_float * 16)(*T.reshape(16).tolist())


def compute_jtj_jtr(J, r):
    """utility::ComputeJTJandJTr<Matrix6f, Vector6f, NumJ> (eigen.inl:120-145) on explicit rows J [n, num_j, 6], r [n, num_j]
    -> (JTJ [6, 6] float32, JTr [6] float32, sum r^2)"""
    Jd = DeviceArray.wrap(J)
    n = Jd.shape[0]
    num_j = int(np.prod(Jd.shape[1:])) // 6
    rd = DeviceArray.wrap(r)
    S = (C.c_double * 32)()
    _lib.check(_lib.lib().cphb_compute_jtj_jtr(Jd.ptr, rd.ptr, n, num_j, S, None))
    return _unpack_sums(S)


def compute_weighted_jtj_jtr(J, r, sigma2, nu):
    """utility::ComputeWeightedJTJandJTr (eigen.inl:147-195) with the RGB-D odometry's Student-t weights
    (odometry.cu:633-648) -> (JTJ, JTr, sum w r^2, w_sum)"""
    Jd = DeviceArray.wrap(J)
    n = Jd.shape[0]
    num_j = int(np.prod(Jd.shape[1:])) // 6
    rd = DeviceArray.wrap(r)
    S = (C.c_double * 32)()
    w = C.c_float(0)
    _lib.check(_lib.lib().cphb_compute_weighted_jtj_jtr(Jd.ptr, rd.ptr, n, num_j, float(sigma2), float(nu), S, C.byref(w), None))
    return _unpack_sums(S) + (float(w.value),)


def _unpack_sums(S):
    S = np.array(S, np.float64)
    JTJ = np.zeros((6, 6), np.float32)
    p = 0
    for a in range(6):
        for b in range(a, 6):
            JTJ[a, b] = JTJ[b, a] = np.float32(S[p])
            p += 1
    return JTJ, S[21:27].astype(np.float32), float(np.float32(S[27]))

