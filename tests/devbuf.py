"""TEST ONLY: raw device buffers through the CUDA runtime (ctypes), for tests that hand device pointers to the C ABI
without importing torch.  libpbk links its own static cudart; both runtimes share the device's primary context, so
pointers are interchangeable."""
import ctypes as C

import numpy as np

_rt = None


def _cudart():
    global _rt
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
