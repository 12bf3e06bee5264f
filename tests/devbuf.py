"""TEST ONLY: raw device buffers through the CUDA runtime (ctypes), for tests that hand device pointers to the C ABI
without importing torch.  libpbk links its own static cudart; both runtimes share the device's primary context, so
pointers are interchangeable."""
import ctypes as C
import os

import numpy as np

_rt = None


class _HostRuntime:
    """PBK_TEST_EMULATED_ABI (tests/conftest.py): "device" memory of the emulated ABI is host memory"""
    def __init__(self):
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        self.libc.free.argtypes = [C.c_void_p]

    def cudaMalloc(self, pp, n):
        pp._obj.value = self.libc.malloc(n)
        return 0 if pp._obj.value else 2

    def cudaFree(self, p):
        self.libc.free(p)
        return 0

    def cudaMemcpy(self, d, s, n, kind):
        C.memmove(d, s, n)
        return 0

    def cudaMemset(self, d, v, n):
        C.memset(d, v, n)
        return 0


def _cudart():
    global _rt
    if _rt is None and os.environ.get("PBK_TEST_EMULATED_ABI"):
        _rt = _HostRuntime()
    if _rt is None:
        last = None
        for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                _rt = C.CDLL(name)
                break
            except OSError as e:
                last = e
        if _rt is None:
            raise last
        _rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        _rt.cudaFree.argtypes = [C.c_void_p]
        _rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
    return _rt


class DevBuf:
    def __init__(self, nbytes: int):
        self.nbytes = max(int(nbytes), 8)
        p = C.c_void_p()
        rc = _cudart().cudaMalloc(C.byref(p), self.nbytes)
        assert rc == 0, f"cudaMalloc({self.nbytes}) -> {rc}"
        self.ptr = p.value
        assert _cudart().cudaMemset(self.ptr, 0, self.nbytes) == 0

    def copy_from(self, other: "DevBuf", dst_off: int, src_off: int, nbytes: int):
        assert dst_off + nbytes <= self.nbytes and src_off + nbytes <= other.nbytes
        assert _cudart().cudaMemcpy(self.ptr + dst_off, other.ptr + src_off, nbytes, 3) == 0       # device to device

    def to_host(self, dtype=np.uint64) -> np.ndarray:
        out = np.zeros(self.nbytes // np.dtype(dtype).itemsize, dtype)
        assert _cudart().cudaMemcpy(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes, 2) == 0
        return out

    def free(self):
        if self.ptr:
            _cudart().cudaFree(self.ptr)
            self.ptr = 0
