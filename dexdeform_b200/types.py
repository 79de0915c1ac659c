"""ctypes binding of libmaniskill_mpm.so -- host-side mirror of the reference's ``mpm/types.py``.

Same public names (``lib``, ``vec3``, ``quat``, ``ivec3``, ``mat3``, ``float32``, ``int32``, ``array``) and the same
argtypes tables (mpm/types.py:103-290) so code written against the reference binding runs unchanged.  Differences:

* the library is built explicitly by ``dexdeform_b200.build`` (nvcc, sm_100a) instead of a JIT ``os.system`` call at
  import (mpm/types.py:12-17); a missing library raises instead of calling ``exit()``;
* ``bind_abi1(lib)`` is a function so that tests can bind the *reference* library with the very same tables.

There is no CPU fallback: every entry point launches CUDA kernels.
"""
import ctypes
import os
from ctypes import c_bool, c_float, c_int32, c_size_t, c_ulonglong, c_void_p

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEXDEFORM_B200_LIB") or os.path.join(PKG_DIR, "libmaniskill_mpm.so")  # env override: A/B-test builds

cuda_stream_t = c_void_p
texture_t = c_ulonglong


class vec3(ctypes.Array):
    _length_ = 3
    _type_ = ctypes.c_float
    length, ctype, size, np_type = 3, ctypes.c_float, 12, np.float32


class quat(ctypes.Array):
    _length_ = 4
    _type_ = ctypes.c_float
    length, ctype, size, np_type = 4, ctypes.c_float, 16, np.float32


class ivec3(ctypes.Array):
    _length_ = 3
    _type_ = ctypes.c_int
    length, ctype, size, np_type = 3, ctypes.c_int, 12, np.int32

    def __repr__(self):
        return f"ivec3({self[0]}, {self[1]}, {self[2]})"


class mat3(ctypes.Array):
    _length_ = 9
    _type_ = ctypes.c_float
    length, ctype, size, np_type = 9, ctypes.c_float, 36, np.float32


class texture_resources(ctypes.Structure):
    _fields_ = [("array", c_void_p), ("texture", texture_t)]


class float32:
    length, ctype, size, np_type = 1, ctypes.c_float, 4, np.float32


class int32:
    length, ctype, size, np_type = 1, ctypes.c_int, 4, np.int32


class int64:
    length, ctype, size, np_type = 1, ctypes.c_int64, 8, np.int64


P = c_void_p
IV = ctypes.POINTER(ivec3)

# symbol -> (restype, argtypes); mirrors mpm/types.py:103-290 line by line
ABI1 = {
    "cuda_alloc": (c_void_p, [c_size_t]),
    "cuda_free": (None, [P]),
    "cuda_upload": (None, [P, P, c_size_t]),
    "cuda_upload_async": (None, [P, P, c_size_t, cuda_stream_t]),
    "cuda_download": (None, [P, P, c_size_t]),
    "cuda_download_async": (None, [P, P, c_size_t, cuda_stream_t]),
    "cuda_copy": (None, [P, P, c_size_t]),
    "cuda_copy_async": (None, [P, P, c_size_t, cuda_stream_t]),
    "cuda_copy2d": (None, [P, c_size_t, P, c_size_t, c_size_t, c_size_t]),
    "cuda_zero": (None, [P, c_size_t]),
    "cuda_zero_async": (None, [P, c_size_t, cuda_stream_t]),
    "cuda_stream_create": (cuda_stream_t, []),
    "cuda_stream_destroy": (None, [cuda_stream_t]),
    "cuda_stream_sync": (None, [cuda_stream_t]),
    "print_memory_info": (None, []),
    "create_volume": (texture_resources, [P, c_int32, c_int32, c_int32]),
    "destroy_volume": (None, [texture_resources]),
    "compute_grid_lower": (None, [P, c_float, c_float, P, c_int32, cuda_stream_t]),
    "compute_svd": (None, [P] * 6 + [c_float, c_int32, cuda_stream_t]),
    "compute_svd_grad": (None, [P] * 11 + [c_float, c_int32, cuda_stream_t]),
    "p2g": (None, [P] * 11 + [IV, c_float, c_float, c_float, P, P, P, c_int32, cuda_stream_t]),
    "p2g_grad": (None, [P] * 11 + [IV, c_float, c_float, c_float] + [P] * 13 + [c_int32, cuda_stream_t]),
    "grid_op_v2": (None, [P] * 11 + [c_float] * 4 + [P, IV, c_int32, cuda_stream_t]),
    "grid_op_v2_grad": (None, [P] * 17 + [c_float] * 4 + [P, P, IV, c_int32, cuda_stream_t]),
    "g2p": (None, [P, P, P, c_float, c_float, c_float, IV, P, c_float, P, P, c_int32, cuda_stream_t]),
    "g2p_grad": (None, [P, P, P, c_float, c_float, c_float, IV, P, c_float, P, P, c_int32] + [P] * 5 + [cuda_stream_t]),
    "render": (None, [P] * 13 + [c_float, IV, c_int32, c_bool, IV, c_int32, c_int32, c_float, c_int32, P, cuda_stream_t]),
    "particle_sdf": (None, [P] * 5 + [c_int32, IV, c_float] + [P] * 3 + [c_int32, cuda_stream_t]),
    "compute_dist": (None, [P] * 6 + [c_int32] + [P] * 4 + [c_int32, c_int32, cuda_stream_t]),
    "particle2mass": (None, [P] * 3 + [IV, c_float, c_float] + [P] * 4 + [c_int32, c_int32, c_int32, cuda_stream_t]),
}


def bind_abi1(library):
    """Attach restype/argtypes of the reference ABI to an already loaded library."""
    for name, (res, args) in ABI1.items():
        fn = getattr(library, name)
        fn.restype = res
        fn.argtypes = args
    return library


def load_library(path=LIB_PATH):
    if not os.path.isfile(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m dexdeform_b200.build` (nvcc, sm_100a). "
            "dexdeform_b200 has no CPU or PyTorch fallback."
        )
    library = bind_abi1(ctypes.cdll.LoadLibrary(path))
    from .engine_abi import bind_abi2
    bind_abi2(library)
    return library


class _LazyLib:
    """Loads the shared object on first attribute access (keeps ``import dexdeform_b200`` cheap and lets the
    build step import the package before the library exists)."""

    _lib = None

    def __getattr__(self, name):
        if _LazyLib._lib is None:
            _LazyLib._lib = load_library()
        return getattr(_LazyLib._lib, name)


lib = _LazyLib()


class array:
    """Device buffer with the interface of mpm/types.py:294-400 (alloc in ctor, zeroed, free in __del__).

    ``library`` selects which loaded library owns the allocation (default: this package's)."""

    def __init__(self, dtype=float32, length=0, library=None):
        assert length > 0
        if dtype == float:
            dtype = float32
        if dtype == int:
            dtype = int32
        self.lib = library if library is not None else lib
        self.dtype = dtype
        self.bytes = dtype.size
        self.nbytes = dtype.size * int(length)
        self.data_ptr = self.lib.cuda_alloc(self.nbytes)
        if not self.data_ptr:
            raise MemoryError(f"cuda_alloc({self.nbytes}) failed")
        self.shape = (int(length),) if dtype.length == 1 else (int(length), dtype.length)
        self.zero()

    def upload_async(self, arr, stream):
        assert arr.shape == self.shape and arr.dtype == self.dtype.np_type and arr.flags["C_CONTIGUOUS"]
        self.lib.cuda_upload_async(self.data_ptr, arr.ctypes.data, self.nbytes, stream)

    def upload(self, arr, strict=False):
        arr = np.ascontiguousarray(arr, dtype=self.dtype.np_type)
        n = arr.shape[0]
        if strict:
            assert n == self.shape[0]
        assert n <= self.shape[0], f"{n} {self.shape[0]}"
        assert (n,) + self.shape[1:] == arr.shape, f"{self.shape}, {arr.shape}"
        self.lib.cuda_upload(self.data_ptr, arr.ctypes.data, self.bytes * n)

    def download(self, n=None, device="numpy", stream=None):
        if device != "numpy":
            import torch
            return torch.tensor(self.download(n, "numpy"), device=device)
        shape = self.shape if n is None else (n,) + self.shape[1:]
        arr = np.empty(shape, self.dtype.np_type)
        self.lib.cuda_download(arr.ctypes.data, self.data_ptr, arr.nbytes)
        return arr

    def cuda_add(self, values, stream=None):
        x = self.download(n=len(values))
        x += np.asarray(values, dtype=x.dtype).reshape(x.shape)
        self.upload(x)

    def zero(self, stream=None):
        if stream is None:
            self.lib.cuda_zero(self.data_ptr, self.nbytes)
        else:
            self.lib.cuda_zero_async(self.data_ptr, self.nbytes, stream)

    def zero_async(self, stream):
        self.lib.cuda_zero_async(self.data_ptr, self.nbytes, stream)

    def __del__(self):
        try:
            if getattr(self, "data_ptr", None):
                self.lib.cuda_free(self.data_ptr)
        except Exception:
            pass

    def __repr__(self):
        return repr(self.download())
