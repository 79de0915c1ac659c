"""Loaders for the checkers (TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and the cpu_baseline /
reference legs of bench.py import this).

``OracleLib``    wraps oracle/libmpm_oracle.so (the C restatement) behind the same method names and argument order as
                 the reference ABI, with host memory standing in for device memory, so one driver can run the product
                 library, the reference library and the oracle on identical buffers.
``load_ref_cpu`` / ``load_ref_gpu`` load the unmodified reference built by oracle/build_ref.sh.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CPU = os.path.join(HERE, "_ref", "libmaniskill_mpm_cpu.so")
REF_GPU = os.path.join(HERE, "_ref", "libmaniskill_mpm.so")
ORACLE = os.path.join(HERE, "libmpm_oracle.so")

_KERNELS = ["compute_svd", "compute_svd_grad", "p2g", "p2g_grad", "grid_op_v2", "grid_op_v2_grad", "g2p", "g2p_grad",
            "compute_dist", "particle2mass"]


def _abi1_tables():
    from dexdeform_b200.types import ABI1, bind_abi1
    return ABI1, bind_abi1


def load_ref_cpu():
    _, bind = _abi1_tables()
    lib = bind(ctypes.cdll.LoadLibrary(REF_CPU))
    lib.ref_cpu_num_threads.restype = ctypes.c_int
    lib.ref_cpu_set_num_threads.argtypes = [ctypes.c_int]
    return lib


def load_ref_gpu():
    _, bind = _abi1_tables()
    return bind(ctypes.cdll.LoadLibrary(REF_GPU))


class OracleLib:
    """The C restatement presented as an ABI-1 'library' operating on host memory."""

    def __init__(self, path=ORACLE):
        abi1, _ = _abi1_tables()
        self._c = ctypes.cdll.LoadLibrary(path)
        self._bufs = {}
        for name in _KERNELS:
            fn = getattr(self._c, "orc_" + name)
            fn.restype = None
            fn.argtypes = abi1[name][1][:-1]  # same order, no stream
            setattr(self, name, (lambda f: (lambda *a: f(*a[:-1])))(fn))
        self._c.orc_num_threads.restype = ctypes.c_int
        self._c.orc_set_num_threads.argtypes = [ctypes.c_int]
        self.raw = self._c

    def num_threads(self):
        return self._c.orc_num_threads()

    def set_num_threads(self, n):
        self._c.orc_set_num_threads(int(n))

    # host stand-ins for the memory helpers
    def cuda_alloc(self, n):
        buf = ctypes.create_string_buffer(int(n))
        ptr = ctypes.addressof(buf)
        self._bufs[ptr] = buf
        return ptr

    def cuda_free(self, ptr):
        self._bufs.pop(ptr, None)

    def cuda_upload(self, d, h, n):
        ctypes.memmove(d, h, n)

    def cuda_download(self, h, d, n):
        ctypes.memmove(h, d, n)

    def cuda_upload_async(self, d, h, n, s):
        ctypes.memmove(d, h, n)

    def cuda_zero(self, p, n):
        ctypes.memset(p, 0, n)

    def cuda_zero_async(self, p, n, s):
        ctypes.memset(p, 0, n)

    def cuda_stream_create(self):
        return None

    def cuda_stream_destroy(self, s):
        pass

    def cuda_stream_sync(self, s):
        pass
