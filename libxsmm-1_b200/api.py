"""Python mirror of the C interface of libxsmm_b200.so (ctypes; no torch, no CPU compute path).

Every function below has the name, argument order and meaning of the reference's C entry point
(reference include/libxsmm_spmdm.h:74-132, include/libxsmm_fsspmdm.h:41-57) or of one of the
stream-ordered additions declared in include/libxsmm_b200.h.  The functions only marshal arguments;
all arithmetic happens in the CUDA kernels inside the shared library.  If the library is missing,
``load()`` raises -- there is no fallback of any kind.

Small conveniences on top (``DeviceBuffer``, ``Spmdm``, ``Fsspmdm``) keep the tests and bench.py short;
they still go through the C ABI for every operation.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libxsmm_b200.so")

LIBXSMM_SPMDM_DATATYPE_F32 = 0
LIBXSMM_SPMDM_DATATYPE_BFLOAT16 = 1

# every symbol include/*.h declares (tests check that the library exports all of them)
REFERENCE_SYMBOLS = [
    "libxsmm_spmdm_init", "libxsmm_spmdm_destroy",
    "libxsmm_spmdm_get_num_createSparseSlice_blocks", "libxsmm_spmdm_get_num_compute_blocks",
    "libxsmm_spmdm_createSparseSlice_fp32_thread", "libxsmm_spmdm_createSparseSlice_bfloat16_thread",
    "libxsmm_spmdm_compute_fp32_thread", "libxsmm_spmdm_compute_bfloat16_thread",
    "libxsmm_dfsspmdm_create", "libxsmm_dfsspmdm_execute", "libxsmm_dfsspmdm_destroy",
    "libxsmm_sfsspmdm_create", "libxsmm_sfsspmdm_execute", "libxsmm_sfsspmdm_destroy",
]
ADDED_SYMBOLS = [
    "libxsmm_spmdm_createSparseSlice_fp32_stream", "libxsmm_spmdm_createSparseSlice_bfloat16_stream",
    "libxsmm_spmdm_compute_fp32_stream", "libxsmm_spmdm_compute_bfloat16_stream",
    "libxsmm_spmdm_exec_host", "libxsmm_spmdm_exec_stream",
    "libxsmm_dfsspmdm_execute_stream", "libxsmm_sfsspmdm_execute_stream",
    "libxsmm_dfsspmdm_is_sparse", "libxsmm_sfsspmdm_is_sparse",
    "libxsmm_dfsspmdm_is_baked", "libxsmm_sfsspmdm_is_baked", "libxsmm_sfsspmdm_is_tensor_core",
    "libxsmm_b200_last_error", "libxsmm_b200_last_error_string", "libxsmm_b200_clear_error",
    "libxsmm_b200_launch_count", "libxsmm_b200_last_compute_kernel",
    "libxsmm_b200_host_alloc", "libxsmm_b200_host_free",
    "libxsmm_b200_device_alloc", "libxsmm_b200_device_free",
    "libxsmm_b200_memcpy_h2d", "libxsmm_b200_memcpy_d2h", "libxsmm_b200_memset",
    "libxsmm_b200_synchronize", "libxsmm_b200_device_count", "libxsmm_b200_set_device",
    "libxsmm_b200_stream_create", "libxsmm_b200_stream_destroy", "libxsmm_b200_stream_synchronize",
    "libxsmm_b200_event_create", "libxsmm_b200_event_destroy", "libxsmm_b200_event_record",
    "libxsmm_b200_event_synchronize", "libxsmm_b200_event_elapsed_ms",
    "libxsmm_b200_memcpy_h2d_async", "libxsmm_b200_memcpy_d2h_async", "libxsmm_b200_memset_async",
    "libxsmm_b200_spmdm_geometry", "libxsmm_b200_fsspmdm_plan", "libxsmm_b200_fsspmdm_kernel_source",
    "libxsmm_b200_free_string",
    "libxsmm_b200_graph_begin", "libxsmm_b200_graph_end", "libxsmm_b200_graph_launch", "libxsmm_b200_graph_destroy",
    "libxsmm_b200_sparse_matmul", "libxsmm_b200_sparse_matmul_cache_entries", "libxsmm_b200_sparse_matmul_cache_clear",
    "libxsmm_b200_csr_read_mtx", "libxsmm_b200_csr_free", "libxsmm_b200_dfsspmdm_create_mtx", "libxsmm_b200_sfsspmdm_create_mtx",
    "libxsmm_b200_dcsr_soa_create", "libxsmm_b200_scsr_soa_create", "libxsmm_b200_csr_soa_execute", "libxsmm_b200_csr_soa_is_baked",
    "libxsmm_b200_csr_soa_destroy", "libxsmm_b200_dcsc_soa_create", "libxsmm_b200_scsc_soa_create",
    "libxsmm_b200_dmmdispatch", "libxsmm_b200_smmdispatch", "libxsmm_b200_mm_execute", "libxsmm_b200_mm_kernel", "libxsmm_b200_mm_release",
]


class libxsmm_spmdm_handle(ctypes.Structure):
    """reference include/libxsmm_spmdm.h:42-60 (same field order; caller-owned storage)."""
    _fields_ = [("m", ctypes.c_int), ("n", ctypes.c_int), ("k", ctypes.c_int),
                ("bm", ctypes.c_int), ("bn", ctypes.c_int), ("bk", ctypes.c_int),
                ("mb", ctypes.c_int), ("nb", ctypes.c_int), ("kb", ctypes.c_int),
                ("datatype", ctypes.c_int),
                ("base_ptr_scratch_A", ctypes.c_void_p),
                ("base_ptr_scratch_B_scratch_C", ctypes.c_void_p),
                ("memory_for_scratch_per_thread", ctypes.c_int)]


class libxsmm_CSR_sparseslice(ctypes.Structure):
    """reference include/libxsmm_spmdm.h:66-71; the three pointers are DEVICE pointers."""
    _fields_ = [("rowidx", ctypes.c_void_p), ("colidx", ctypes.c_void_p), ("values", ctypes.c_void_p)]


_lib = None


def load():
    """dlopen the product library.  Raises if it has not been built (python libxsmm-1_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libxsmm_b200.so is not built (%s): run `python libxsmm-1_b200/build.py`; "
                           "there is no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cc = ctypes.c_void_p, ctypes.c_int, ctypes.c_char
    H = ctypes.POINTER(libxsmm_spmdm_handle)
    S = ctypes.POINTER(libxsmm_CSR_sparseslice)
    L.libxsmm_spmdm_init.argtypes = [ci, ci, ci, ci, H, ctypes.POINTER(S)]
    L.libxsmm_spmdm_init.restype = None
    L.libxsmm_spmdm_destroy.argtypes = [H]
    L.libxsmm_spmdm_destroy.restype = None
    L.libxsmm_spmdm_get_num_createSparseSlice_blocks.argtypes = [H]
    L.libxsmm_spmdm_get_num_compute_blocks.argtypes = [H]
    for nm in ("fp32", "bfloat16"):
        f = getattr(L, "libxsmm_spmdm_createSparseSlice_%s_thread" % nm)
        f.argtypes = [H, cc, vp, S, ci, ci, ci]
        f.restype = None
        f = getattr(L, "libxsmm_spmdm_compute_%s_thread" % nm)
        f.argtypes = [H, cc, cc, vp, S, vp, cc, vp, vp, ci, ci, ci]
        f.restype = None
        f = getattr(L, "libxsmm_spmdm_createSparseSlice_%s_stream" % nm)
        f.argtypes = [H, cc, vp, S, vp]
        f.restype = None
        f = getattr(L, "libxsmm_spmdm_compute_%s_stream" % nm)
        f.argtypes = [H, cc, cc, vp, S, vp, cc, vp, vp, vp]
        f.restype = None
    L.libxsmm_spmdm_exec_host.argtypes = [H, S, ci, cc, cc, cc, vp, vp, vp, vp]
    L.libxsmm_spmdm_exec_host.restype = None
    L.libxsmm_spmdm_exec_stream.argtypes = [H, S, ci, cc, cc, cc, vp, vp, vp, vp, vp]
    L.libxsmm_spmdm_exec_stream.restype = None
    L.libxsmm_dfsspmdm_create.argtypes = [ci] * 6 + [ctypes.c_double, ctypes.c_double, vp]
    L.libxsmm_dfsspmdm_create.restype = vp
    L.libxsmm_sfsspmdm_create.argtypes = [ci] * 6 + [ctypes.c_float, ctypes.c_float, vp]
    L.libxsmm_sfsspmdm_create.restype = vp
    for p in ("d", "s"):
        f = getattr(L, "libxsmm_%sfsspmdm_execute" % p)
        f.argtypes = [vp, vp, vp]
        f.restype = None
        f = getattr(L, "libxsmm_%sfsspmdm_execute_stream" % p)
        f.argtypes = [vp, vp, vp, vp]
        f.restype = None
        f = getattr(L, "libxsmm_%sfsspmdm_destroy" % p)
        f.argtypes = [vp]
        f.restype = None
        getattr(L, "libxsmm_%sfsspmdm_is_sparse" % p).argtypes = [vp]
        getattr(L, "libxsmm_%sfsspmdm_is_baked" % p).argtypes = [vp]
    L.libxsmm_sfsspmdm_is_tensor_core.argtypes = [vp]
    L.libxsmm_b200_last_error_string.restype = ctypes.c_char_p
    L.libxsmm_b200_launch_count.restype = ctypes.c_ulonglong
    L.libxsmm_b200_last_compute_kernel.restype = ctypes.c_char_p
    L.libxsmm_b200_host_alloc.argtypes = [ctypes.c_size_t]
    L.libxsmm_b200_host_alloc.restype = vp
    L.libxsmm_b200_host_free.argtypes = [vp]
    L.libxsmm_b200_device_alloc.argtypes = [ctypes.c_size_t]
    L.libxsmm_b200_device_alloc.restype = vp
    L.libxsmm_b200_device_free.argtypes = [vp]
    L.libxsmm_b200_memcpy_h2d.argtypes = [vp, vp, ctypes.c_size_t]
    L.libxsmm_b200_memcpy_d2h.argtypes = [vp, vp, ctypes.c_size_t]
    L.libxsmm_b200_memset.argtypes = [vp, ci, ctypes.c_size_t]
    L.libxsmm_b200_set_device.argtypes = [ci]
    L.libxsmm_b200_stream_create.restype = vp
    L.libxsmm_b200_stream_destroy.argtypes = [vp]
    L.libxsmm_b200_stream_synchronize.argtypes = [vp]
    L.libxsmm_b200_event_create.restype = vp
    L.libxsmm_b200_event_destroy.argtypes = [vp]
    L.libxsmm_b200_event_record.argtypes = [vp, vp]
    L.libxsmm_b200_event_synchronize.argtypes = [vp]
    L.libxsmm_b200_event_elapsed_ms.argtypes = [vp, vp]
    L.libxsmm_b200_event_elapsed_ms.restype = ctypes.c_float
    L.libxsmm_b200_memcpy_h2d_async.argtypes = [vp, vp, ctypes.c_size_t, vp]
    L.libxsmm_b200_memcpy_d2h_async.argtypes = [vp, vp, ctypes.c_size_t, vp]
    L.libxsmm_b200_memset_async.argtypes = [vp, ci, ctypes.c_size_t, vp]
    L.libxsmm_b200_graph_begin.argtypes = [vp]
    L.libxsmm_b200_graph_end.argtypes = [vp]
    L.libxsmm_b200_graph_end.restype = vp
    L.libxsmm_b200_graph_launch.argtypes = [vp, vp]
    L.libxsmm_b200_graph_destroy.argtypes = [vp]
    L.libxsmm_b200_spmdm_geometry.argtypes = [ci] * 5 + [vp]
    L.libxsmm_b200_fsspmdm_plan.argtypes = [ci] * 7 + [ctypes.c_double, vp, vp]
    L.libxsmm_b200_fsspmdm_kernel_source.argtypes = [ci] * 7 + [ctypes.c_double, vp]
    L.libxsmm_b200_fsspmdm_kernel_source.restype = vp
    L.libxsmm_b200_free_string.argtypes = [vp]
    _lib = L
    return L


def device_count():
    return int(load().libxsmm_b200_device_count())


def require_gpu():
    if device_count() < 1:
        raise RuntimeError("libxsmm_b200: no CUDA device visible; this library has no CPU path")


def last_error():
    L = load()
    return int(L.libxsmm_b200_last_error()), (L.libxsmm_b200_last_error_string() or b"").decode()


def clear_error():
    load().libxsmm_b200_clear_error()


def check():
    code, msg = last_error()
    if code != 0:
        clear_error()
        raise RuntimeError("libxsmm_b200 error %d: %s" % (code, msg))


def launch_count():
    return int(load().libxsmm_b200_launch_count())


def last_compute_kernel():
    return load().libxsmm_b200_last_compute_kernel().decode()


def synchronize():
    load().libxsmm_b200_synchronize()
    check()


def _addr(x):
    """numpy array -> host address; DeviceBuffer/HostBuffer -> its address; int passes through."""
    if x is None:
        return None
    if isinstance(x, (DeviceBuffer, HostBuffer)):
        return x.ptr
    if isinstance(x, np.ndarray):
        assert x.flags.c_contiguous
        return x.ctypes.data
    return int(x)


class DeviceBuffer:
    """cudaMalloc'ed bytes obtained through the C ABI (libxsmm_b200_device_alloc)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = load().libxsmm_b200_device_alloc(self.nbytes)
        if not self.ptr:
            check()
            raise MemoryError("device_alloc(%d)" % self.nbytes)

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        d = cls(max(a.nbytes, 1))
        if a.nbytes:
            load().libxsmm_b200_memcpy_h2d(d.ptr, a.ctypes.data, a.nbytes)
        check()
        return d

    def upload(self, a, offset=0):
        a = np.ascontiguousarray(a)
        assert offset + a.nbytes <= self.nbytes
        load().libxsmm_b200_memcpy_h2d(self.ptr + offset, a.ctypes.data, a.nbytes)
        check()

    def to_numpy(self, dtype, shape, offset=0):
        out = np.empty(shape, dtype)
        assert offset + out.nbytes <= self.nbytes
        if out.nbytes:
            load().libxsmm_b200_memcpy_d2h(out.ctypes.data, self.ptr + offset, out.nbytes)
        check()
        return out

    def fill(self, byte):
        load().libxsmm_b200_memset(self.ptr, int(byte), self.nbytes)

    def free(self):
        if self.ptr:
            load().libxsmm_b200_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class HostBuffer:
    """Page-locked host memory (libxsmm_b200_host_alloc) viewed as a numpy array."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = load().libxsmm_b200_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            check()
            raise MemoryError("host_alloc(%d)" % self.nbytes)
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            load().libxsmm_b200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Stream:
    def __init__(self):
        self.ptr = load().libxsmm_b200_stream_create()
        check()

    def synchronize(self):
        load().libxsmm_b200_stream_synchronize(self.ptr)
        check()

    def destroy(self):
        if self.ptr:
            load().libxsmm_b200_stream_destroy(self.ptr)
            self.ptr = None


class Event:
    def __init__(self):
        self.ptr = load().libxsmm_b200_event_create()
        check()

    def record(self, stream):
        load().libxsmm_b200_event_record(self.ptr, stream.ptr if isinstance(stream, Stream) else stream)

    def synchronize(self):
        load().libxsmm_b200_event_synchronize(self.ptr)

    def elapsed_ms(self, later):
        return float(load().libxsmm_b200_event_elapsed_ms(self.ptr, later.ptr))

    def destroy(self):
        if self.ptr:
            load().libxsmm_b200_event_destroy(self.ptr)
            self.ptr = None


def _c(ch):
    return ch.encode() if isinstance(ch, str) else ch


def _sptr(stream):
    return stream.ptr if isinstance(stream, Stream) else stream


# ======================================================================================================
# SPMDM -- same names and argument order as reference include/libxsmm_spmdm.h:74-132
# ======================================================================================================
def libxsmm_spmdm_init(M, N, K, max_threads):
    """-> (handle, slices).  handle is caller storage filled by the library; slices points INTO the
    library-owned arena (reference src/libxsmm_spmdm.c:540-627)."""
    L = load()
    h = libxsmm_spmdm_handle()
    s = ctypes.POINTER(libxsmm_CSR_sparseslice)()
    L.libxsmm_spmdm_init(M, N, K, max_threads, ctypes.byref(h), ctypes.byref(s))
    return h, s


def libxsmm_spmdm_destroy(handle):
    load().libxsmm_spmdm_destroy(ctypes.byref(handle))


def libxsmm_spmdm_get_num_createSparseSlice_blocks(handle):
    return int(load().libxsmm_spmdm_get_num_createSparseSlice_blocks(ctypes.byref(handle)))


def libxsmm_spmdm_get_num_compute_blocks(handle):
    return int(load().libxsmm_spmdm_get_num_compute_blocks(ctypes.byref(handle)))


def libxsmm_spmdm_createSparseSlice_fp32_thread(handle, transa, a, slices, block_id, tid, nthreads):
    load().libxsmm_spmdm_createSparseSlice_fp32_thread(ctypes.byref(handle), _c(transa), _addr(a), slices, block_id, tid, nthreads)


def libxsmm_spmdm_createSparseSlice_bfloat16_thread(handle, transa, a, slices, block_id, tid, nthreads):
    load().libxsmm_spmdm_createSparseSlice_bfloat16_thread(ctypes.byref(handle), _c(transa), _addr(a), slices, block_id, tid, nthreads)


def libxsmm_spmdm_compute_fp32_thread(handle, transa, transb, alpha, slices, b, transc, beta, c, block_id, tid, nthreads):
    al = ctypes.c_float(alpha)
    be = ctypes.c_float(beta)
    load().libxsmm_spmdm_compute_fp32_thread(ctypes.byref(handle), _c(transa), _c(transb), ctypes.addressof(al), slices,
                                             _addr(b), _c(transc), ctypes.addressof(be), _addr(c), block_id, tid, nthreads)


def libxsmm_spmdm_compute_bfloat16_thread(handle, transa, transb, alpha_bits, slices, b, transc, beta_bits, c, block_id, tid, nthreads):
    """alpha_bits / beta_bits are raw 16-bit patterns, read by the library the way the reference reads
    them (as integers; reference bf16 compute template :91,113,164)."""
    al = ctypes.c_ushort(alpha_bits)
    be = ctypes.c_ushort(beta_bits)
    load().libxsmm_spmdm_compute_bfloat16_thread(ctypes.byref(handle), _c(transa), _c(transb), ctypes.addressof(al), slices,
                                                 _addr(b), _c(transc), ctypes.addressof(be), _addr(c), block_id, tid, nthreads)


# stream-ordered additions (include/libxsmm_b200.h)
def libxsmm_spmdm_createSparseSlice_fp32_stream(handle, transa, d_a, slices, stream=None):
    load().libxsmm_spmdm_createSparseSlice_fp32_stream(ctypes.byref(handle), _c(transa), _addr(d_a), slices, _sptr(stream))


def libxsmm_spmdm_createSparseSlice_bfloat16_stream(handle, transa, d_a, slices, stream=None):
    load().libxsmm_spmdm_createSparseSlice_bfloat16_stream(ctypes.byref(handle), _c(transa), _addr(d_a), slices, _sptr(stream))


def libxsmm_spmdm_compute_fp32_stream(handle, transa, transb, alpha, slices, d_b, transc, beta, d_c, stream=None):
    al = ctypes.c_float(alpha)
    be = ctypes.c_float(beta)
    load().libxsmm_spmdm_compute_fp32_stream(ctypes.byref(handle), _c(transa), _c(transb), ctypes.addressof(al), slices,
                                             _addr(d_b), _c(transc), ctypes.addressof(be), _addr(d_c), _sptr(stream))


def libxsmm_spmdm_compute_bfloat16_stream(handle, transa, transb, alpha_bits, slices, d_b, transc, beta_bits, d_c, stream=None):
    al = ctypes.c_ushort(alpha_bits)
    be = ctypes.c_ushort(beta_bits)
    load().libxsmm_spmdm_compute_bfloat16_stream(ctypes.byref(handle), _c(transa), _c(transb), ctypes.addressof(al), slices,
                                                 _addr(d_b), _c(transc), ctypes.addressof(be), _addr(d_c), _sptr(stream))


def _beta_box(datatype, beta):
    return ctypes.c_ushort(int(beta)) if datatype == LIBXSMM_SPMDM_DATATYPE_BFLOAT16 else ctypes.c_float(beta)


def libxsmm_spmdm_exec_host(handle, slices, datatype, transa, transb, transc, a, b, beta, c):
    """One whole multiply on HOST matrices (what one repetition of samples/spmdm/spmdm.c:88-111 does)."""
    be = _beta_box(datatype, beta)
    load().libxsmm_spmdm_exec_host(ctypes.byref(handle), slices, datatype, _c(transa), _c(transb), _c(transc),
                                   _addr(a), _addr(b), ctypes.addressof(be), _addr(c))


def libxsmm_spmdm_exec_stream(handle, slices, datatype, transa, transb, transc, d_a, d_b, beta, d_c, stream=None):
    """Slice creation followed by compute, both asynchronous on ``stream``, DEVICE matrices."""
    be = _beta_box(datatype, beta)
    load().libxsmm_spmdm_exec_stream(ctypes.byref(handle), slices, datatype, _c(transa), _c(transb), _c(transc),
                                     _addr(d_a), _addr(d_b), ctypes.addressof(be), _addr(d_c), _sptr(stream))


def libxsmm_b200_sparse_matmul(datatype, transa, transb, transc, M, N, K, max_threads, d_a, d_b, beta, d_c, stream=None):
    """Fused caller step (handle cache + slice creation + compute in one stream-ordered call); returns the error code."""
    be = _beta_box(datatype, beta)
    L = load()
    L.libxsmm_b200_sparse_matmul.restype = ctypes.c_int
    L.libxsmm_b200_sparse_matmul.argtypes = [ctypes.c_int, ctypes.c_char, ctypes.c_char, ctypes.c_char, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return int(L.libxsmm_b200_sparse_matmul(datatype, _c(transa), _c(transb), _c(transc), M, N, K, max_threads,
                                            _addr(d_a), _addr(d_b), ctypes.addressof(be), _addr(d_c), _sptr(stream)))


def sparse_matmul_cache_entries():
    return int(load().libxsmm_b200_sparse_matmul_cache_entries())


def sparse_matmul_cache_clear():
    load().libxsmm_b200_sparse_matmul_cache_clear()


def csr_read_mtx(path):
    """MatrixMarket coordinate file -> (row_ptr, col_idx, values, rows, cols) through the library's reader (host only).
    Raises ValueError with the library's message on a malformed file."""
    L = load()
    rp, ci, va = ctypes.POINTER(ctypes.c_uint)(), ctypes.POINTER(ctypes.c_uint)(), ctypes.POINTER(ctypes.c_double)()
    nr, nc, ne = ctypes.c_uint(0), ctypes.c_uint(0), ctypes.c_uint(0)
    L.libxsmm_b200_csr_read_mtx.restype = ctypes.c_int
    rc = L.libxsmm_b200_csr_read_mtx(os.fsencode(path), ctypes.byref(rp), ctypes.byref(ci), ctypes.byref(va),
                                     ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(ne))
    if 0 != rc:
        code, msg = last_error()
        clear_error()
        raise ValueError("csr_read_mtx failed (%d): %s" % (rc, msg))
    try:
        row_ptr = np.ctypeslib.as_array(rp, shape=(nr.value + 1,)).copy()
        col_idx = np.ctypeslib.as_array(ci, shape=(ne.value,)).copy()
        values = np.ctypeslib.as_array(va, shape=(ne.value,)).copy()
    finally:
        L.libxsmm_b200_csr_free.restype = None
        L.libxsmm_b200_csr_free(rp, ci, va)
    return row_ptr, col_idx, values, int(nr.value), int(nc.value)


# ======================================================================================================
# FSSPMDM -- reference include/libxsmm_fsspmdm.h:41-57
# ======================================================================================================
def libxsmm_dfsspmdm_create(M, N, K, lda, ldb, ldc, alpha, beta, a_dense):
    a = np.ascontiguousarray(a_dense, np.float64)
    return load().libxsmm_dfsspmdm_create(M, N, K, lda, ldb, ldc, float(alpha), float(beta), a.ctypes.data)


def libxsmm_sfsspmdm_create(M, N, K, lda, ldb, ldc, alpha, beta, a_dense):
    a = np.ascontiguousarray(a_dense, np.float32)
    return load().libxsmm_sfsspmdm_create(M, N, K, lda, ldb, ldc, float(alpha), float(beta), a.ctypes.data)


def libxsmm_dfsspmdm_execute(handle, B, C):
    load().libxsmm_dfsspmdm_execute(handle, _addr(B), _addr(C))


def libxsmm_sfsspmdm_execute(handle, B, C):
    load().libxsmm_sfsspmdm_execute(handle, _addr(B), _addr(C))


def libxsmm_dfsspmdm_execute_stream(handle, d_B, d_C, stream=None):
    load().libxsmm_dfsspmdm_execute_stream(handle, _addr(d_B), _addr(d_C), _sptr(stream))


def libxsmm_sfsspmdm_execute_stream(handle, d_B, d_C, stream=None):
    load().libxsmm_sfsspmdm_execute_stream(handle, _addr(d_B), _addr(d_C), _sptr(stream))


def libxsmm_dfsspmdm_destroy(handle):
    load().libxsmm_dfsspmdm_destroy(handle)


def libxsmm_sfsspmdm_destroy(handle):
    load().libxsmm_sfsspmdm_destroy(handle)


# ======================================================================================================
# conveniences used by tests and bench.py
# ======================================================================================================
class Spmdm:
    """One spmdm problem: handle + slices, with helpers to read the slice arena back in the flat layout
    the oracle uses (slice s = kb*mb + mb owns rowidx[s, :bm+1], colidx[s, :bm*bk], values[s, :bm*bk])."""

    def __init__(self, M, N, K, max_threads=1):
        require_gpu()
        self.handle, self.slices = libxsmm_spmdm_init(M, N, K, max_threads)
        check()
        if not self.handle.base_ptr_scratch_A:
            raise RuntimeError("libxsmm_spmdm_init failed")

    @property
    def geometry(self):
        h = self.handle
        return dict(m=h.m, n=h.n, k=h.k, bm=h.bm, bn=h.bn, bk=h.bk, mb=h.mb, nb=h.nb, kb=h.kb)

    def create_slices(self, d_a, transa="N", bf16=False, stream=None):
        f = libxsmm_spmdm_createSparseSlice_bfloat16_stream if bf16 else libxsmm_spmdm_createSparseSlice_fp32_stream
        f(self.handle, transa, d_a, self.slices, stream)

    def compute(self, d_b, d_c, transb="N", transc="N", beta=0.0, bf16=False, stream=None, transa="N"):
        if bf16:
            libxsmm_spmdm_compute_bfloat16_stream(self.handle, transa, transb, 0x3F80, self.slices, d_b, transc, int(beta), d_c, stream)
        else:
            libxsmm_spmdm_compute_fp32_stream(self.handle, transa, transb, 1.0, self.slices, d_b, transc, float(beta), d_c, stream)

    def read_slices(self):
        h = self.handle
        ns, cap = h.mb * h.kb, h.bm * h.bk
        L = load()
        L.libxsmm_b200_synchronize()
        ro = np.empty((ns, h.bm + 1), np.uint16)
        co = np.empty((ns, cap), np.uint16)
        va = np.empty((ns, cap), np.float32)
        # the table entries point into one arena with per-array strides; slice 0 is the base of each
        s0 = self.slices[0]
        L.libxsmm_b200_memcpy_d2h(ro.ctypes.data, s0.rowidx, ro.nbytes)
        L.libxsmm_b200_memcpy_d2h(co.ctypes.data, s0.colidx, co.nbytes)
        L.libxsmm_b200_memcpy_d2h(va.ctypes.data, s0.values, va.nbytes)
        check()
        return ro, co, va

    def destroy(self):
        if self.handle.base_ptr_scratch_A:
            libxsmm_spmdm_destroy(self.handle)


class Fsspmdm:
    """A fixed operator: create once, execute many times (reference src/libxsmm_fsspmdm.c)."""

    def __init__(self, a_dense, N, ldb=None, ldc=None, beta=0.0, alpha=1.0, lda=None):
        require_gpu()
        a = np.ascontiguousarray(a_dense)
        assert a.dtype in (np.float64, np.float32)
        self.double = a.dtype == np.float64
        self.M, self.K = a.shape
        self.N = N
        self.ldb = N if ldb is None else ldb
        self.ldc = N if ldc is None else ldc
        lda = self.K if lda is None else lda
        create = libxsmm_dfsspmdm_create if self.double else libxsmm_sfsspmdm_create
        self.handle = create(self.M, N, self.K, lda, self.ldb, self.ldc, alpha, beta, a)
        if not self.handle:
            code, msg = last_error()
            clear_error()
            raise ValueError("fsspmdm_create failed: %s" % msg)
        clear_error()   # a missing NVRTC is reported but not fatal (generic kernel)

    @classmethod
    def from_mtx(cls, path, N, ldb=None, ldc=None, beta=0.0, double=True):
        """the operator of a MatrixMarket file (libxsmm_b200_[sd]fsspmdm_create_mtx)."""
        require_gpu()
        L = load()
        self = cls.__new__(cls)
        self.double, self.N = bool(double), N
        self.ldb = N if ldb is None else ldb
        self.ldc = N if ldc is None else ldc
        m, k = ctypes.c_int(0), ctypes.c_int(0)
        f = L.libxsmm_b200_dfsspmdm_create_mtx if double else L.libxsmm_b200_sfsspmdm_create_mtx
        f.restype = ctypes.c_void_p
        f.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double if double else ctypes.c_float,
                      ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        self.handle = f(os.fsencode(path), N, self.ldb, self.ldc, float(beta), ctypes.byref(m), ctypes.byref(k))
        if not self.handle:
            code, msg = last_error()
            clear_error()
            raise ValueError("fsspmdm_create_mtx failed: %s" % msg)
        clear_error()
        self.M, self.K = int(m.value), int(k.value)
        return self

    @property
    def is_sparse(self):
        L = load()
        return bool((L.libxsmm_dfsspmdm_is_sparse if self.double else L.libxsmm_sfsspmdm_is_sparse)(self.handle))

    @property
    def is_baked(self):
        L = load()
        return bool((L.libxsmm_dfsspmdm_is_baked if self.double else L.libxsmm_sfsspmdm_is_baked)(self.handle))

    @property
    def is_tensor_core(self):
        return (not self.double) and bool(load().libxsmm_sfsspmdm_is_tensor_core(self.handle))

    def execute(self, B, C):
        (libxsmm_dfsspmdm_execute if self.double else libxsmm_sfsspmdm_execute)(self.handle, B, C)

    def execute_stream(self, d_B, d_C, stream=None):
        (libxsmm_dfsspmdm_execute_stream if self.double else libxsmm_sfsspmdm_execute_stream)(self.handle, d_B, d_C, stream)

    def destroy(self):
        if self.handle:
            (libxsmm_dfsspmdm_destroy if self.double else libxsmm_sfsspmdm_destroy)(self.handle)
            self.handle = None


class MmDispatch:
    """libxsmm_b200_[sd]mmdispatch: column-major SMM handle used the way samples/pyfr/pyfr_gemm_rm.c:98-122 uses
    libxsmm_dmmdispatch (row-major operator applied to a panel with very many columns)."""

    def __init__(self, m, n, k, lda=None, ldb=None, ldc=None, alpha=None, beta=None, dtype=np.float64):
        require_gpu()
        L = load()
        self.double = np.dtype(dtype) == np.float64
        f = L.libxsmm_b200_dmmdispatch if self.double else L.libxsmm_b200_smmdispatch
        f.restype = ctypes.c_void_p
        f.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5
        sc = ctypes.c_double if self.double else ctypes.c_float
        box = lambda v, t: None if v is None else ctypes.cast(ctypes.pointer(t(v)), ctypes.c_void_p)   # noqa: E731
        self.handle = f(m, n, k, box(lda, ctypes.c_int), box(ldb, ctypes.c_int), box(ldc, ctypes.c_int), box(alpha, sc), box(beta, sc))
        if not self.handle:
            code, msg = last_error()
            clear_error()
            raise ValueError("mmdispatch failed: %s" % msg)

    def execute(self, d_a, b, d_c, m_total, stream=None):
        L = load()
        L.libxsmm_b200_mm_execute.restype = None
        L.libxsmm_b200_mm_execute.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
        L.libxsmm_b200_mm_execute(self.handle, _addr(d_a), _addr(b), _addr(d_c), m_total, _sptr(stream))

    @property
    def kernel(self):
        L = load()
        L.libxsmm_b200_mm_kernel.restype = ctypes.c_char_p
        L.libxsmm_b200_mm_kernel.argtypes = [ctypes.c_void_p]
        return L.libxsmm_b200_mm_kernel(self.handle).decode()

    def release(self):
        if self.handle:
            L = load()
            L.libxsmm_b200_mm_release.argtypes = [ctypes.c_void_p]
            L.libxsmm_b200_mm_release(self.handle)
            self.handle = None


class CsrSoa:
    """CSR operand applied to [element][row][column][soa] tensors (libxsmm_b200_[sd]csr_soa_*; the batched GPU counterpart of
    libxsmm_create_xcsr_soa).  sparse="A": reference samples/edge/asparse_srsoa.c (the CSR matrix is M x K, the dense operand
    B is [K][ldb][soa]); sparse="B": samples/edge/bsparse_srsoa.c (the CSR matrix is K x N, the dense operand A is [M][lda][soa])."""

    def __init__(self, M, N, K, rowptr, colidx, values, soa_width, lda=None, ldb=None, ldc=None, beta=0.0, sparse="A", fmt="csr"):
        """fmt="csc" (sparse="B" only): rowptr / colidx are the column pointers / row indices of B (libxsmm_create_xcsc_soa,
        reference samples/edge/bsparse_scsoa.c)."""
        require_gpu()
        L = load()
        values = np.ascontiguousarray(values)
        assert values.dtype in (np.float64, np.float32) and sparse in ("A", "B") and fmt in ("csr", "csc") and (fmt == "csr" or sparse == "B")
        self.double = values.dtype == np.float64
        rowptr = np.ascontiguousarray(rowptr, np.uint32); colidx = np.ascontiguousarray(colidx, np.uint32)
        self.M, self.N, self.K, self.soa, self.sparse = M, N, K, soa_width, sparse
        self.lda = 0 if sparse == "A" else (K if lda is None else lda)
        self.ldb = 0 if sparse == "B" else (N if ldb is None else ldb)
        self.ldc = N if ldc is None else ldc
        if fmt == "csc":
            f = L.libxsmm_b200_dcsc_soa_create if self.double else L.libxsmm_b200_scsc_soa_create
            f.restype = ctypes.c_void_p
            f.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double if self.double else ctypes.c_float] + [ctypes.c_void_p] * 3
            self.handle = f(M, N, K, self.lda, self.ldc, soa_width, float(beta), rowptr.ctypes.data, colidx.ctypes.data, values.ctypes.data)
        else:
            f = L.libxsmm_b200_dcsr_soa_create if self.double else L.libxsmm_b200_scsr_soa_create
            f.restype = ctypes.c_void_p
            f.argtypes = [ctypes.c_int] * 7 + [ctypes.c_double if self.double else ctypes.c_float] + [ctypes.c_void_p] * 3
            self.handle = f(M, N, K, self.lda, self.ldb, self.ldc, soa_width, float(beta), rowptr.ctypes.data, colidx.ctypes.data, values.ctypes.data)
        if not self.handle:
            code, msg = last_error()
            clear_error()
            raise ValueError("csr_soa_create failed: %s" % msg)
        clear_error()

    @property
    def is_baked(self):
        return bool(load().libxsmm_b200_csr_soa_is_baked(ctypes.c_void_p(self.handle)))

    def execute(self, d_X, d_C, n_elements, stride_x=None, stride_c=None, stream=None):
        L = load()
        L.libxsmm_b200_csr_soa_execute.restype = None
        L.libxsmm_b200_csr_soa_execute.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p]
        dense = self.K * self.ldb if self.sparse == "A" else self.M * self.lda
        sx = dense * self.soa if stride_x is None else stride_x
        sc = self.M * self.ldc * self.soa if stride_c is None else stride_c
        L.libxsmm_b200_csr_soa_execute(self.handle, _addr(d_X), _addr(d_C), n_elements, sx, sc, _sptr(stream))

    def destroy(self):
        if self.handle:
            L = load()
            L.libxsmm_b200_csr_soa_destroy.argtypes = [ctypes.c_void_p]
            L.libxsmm_b200_csr_soa_destroy(self.handle)
            self.handle = None


# ---- host-only planning (works without a GPU) ----------------------------------------------------------
def spmdm_geometry(M, N, K, max_threads=1, bn=48):
    vec = (ctypes.c_int * 9)()
    if 0 != load().libxsmm_b200_spmdm_geometry(M, N, K, max_threads, bn, ctypes.cast(vec, ctypes.c_void_p)):
        raise ValueError("spmdm_geometry(%d,%d,%d)" % (M, N, K))
    return dict(zip(("m", "n", "k", "bm", "bn", "bk", "mb", "nb", "kb"), [int(v) for v in vec]))


def fsspmdm_plan(a_dense, N=16, ldb=None, ldc=None, beta=0.0, lda=None):
    a = np.ascontiguousarray(a_dense)
    assert a.dtype in (np.float64, np.float32)
    M, K = a.shape
    info = (ctypes.c_longlong * 6)()
    rc = load().libxsmm_b200_fsspmdm_plan(int(a.dtype == np.float64), M, N, K, K if lda is None else lda,
                                          N if ldb is None else ldb, N if ldc is None else ldc, float(beta),
                                          a.ctypes.data, ctypes.cast(info, ctypes.c_void_p))
    if rc != 0:
        code, msg = last_error()
        clear_error()
        raise ValueError(msg)
    return dict(nnz=int(info[0]), n_unique=int(info[1]), sparse=bool(info[2]), x86_code_size=int(info[3]), chunk=int(info[4]),
                form={0: "generic", 1: "baked-registers", 2: "baked-strip"}[int(info[5])])


def fsspmdm_kernel_source(a_dense, N=16, ldb=None, ldc=None, beta=0.0):
    a = np.ascontiguousarray(a_dense)
    M, K = a.shape
    L = load()
    p = L.libxsmm_b200_fsspmdm_kernel_source(int(a.dtype == np.float64), M, N, K, K, N if ldb is None else ldb,
                                             N if ldc is None else ldc, float(beta), a.ctypes.data)
    if not p:
        return None
    src = ctypes.string_at(p).decode()
    L.libxsmm_b200_free_string(p)
    return src
