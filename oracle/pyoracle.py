"""TEST INFRASTRUCTURE ONLY: ctypes front-ends for the checker libraries.

* ``Oracle``  -> oracle/liboracle.so, the CPU restatement (oracle/*.c).
* ``Ref``     -> oracle/_ref/libref_driver.so, the UNMODIFIED reference compiled by
                 oracle/build_ref.sh and driven the way its samples drive it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_c_int_p = ctypes.POINTER(ctypes.c_int)


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def simd_width_for_bn(bn):
    """bn -> vector width of the reference instantiation (src/libxsmm_spmdm.c:557-583)."""
    return {96: 16, 48: 8}.get(int(bn), 1)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])


def build_ref(flavor="avx2"):
    subprocess.check_call([os.path.join(HERE, "build_ref.sh"), flavor],
                          stdout=subprocess.DEVNULL)


class Geometry(dict):
    """m n k bm bn bk mb nb kb (+ scratch) as attributes; .vec is the int[9] the C side wants."""

    def __getattr__(self, k):
        return self[k]

    @property
    def vec(self):
        return (ctypes.c_int * 9)(*[self[k] for k in ("m", "n", "k", "bm", "bn", "bk", "mb", "nb", "kb")])

    @property
    def nslices(self):
        return self["mb"] * self["kb"]

    @property
    def cap(self):
        return self["bm"] * self["bk"]


def _geom_from(vec, scratch):
    g = Geometry(zip(("m", "n", "k", "bm", "bn", "bk", "mb", "nb", "kb"), [int(v) for v in vec]))
    g["scratch"] = int(scratch)
    return g


def empty_slices(g):
    ns = g.nslices
    return (np.zeros((ns, g.bm + 1), np.uint16), np.zeros((ns, g.cap), np.uint16),
            np.zeros((ns, g.cap), np.float32))


def slice_counts(g, rowidx):
    """true nnz per slice (row pointers are u16 and may have wrapped)."""
    out = np.zeros(g.nslices, np.int64)
    for s in range(g.nslices):
        mb = s % g.mb
        nrows = min(g.bm, g.m - mb * g.bm)
        d = (rowidx[s, 1:nrows + 1].astype(np.int64) - rowidx[s, :nrows].astype(np.int64)) & 0xFFFF
        out[s] = d.sum()
    return out


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = L = ctypes.CDLL(path)
        L.orc_spmdm_geometry.restype = ctypes.c_int
        L.orc_spmdm_compute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char, ctypes.c_char,
                                        ctypes.c_float] + [ctypes.c_void_p] * 5 + [ctypes.c_int] * 2
        L.orc_spmdm_slices.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char, ctypes.c_void_p,
                                       ctypes.c_int] + [ctypes.c_void_p] * 3
        L.orc_dfsspmdm_execute.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double] + [ctypes.c_void_p] * 3 + [ctypes.c_int]
        L.orc_sfsspmdm_execute.argtypes = [ctypes.c_int] * 6 + [ctypes.c_float] + [ctypes.c_void_p] * 3
        L.orc_dfsspmdm_branch.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] + [ctypes.c_int] * 2 + [ctypes.c_double, ctypes.c_int]
        L.orc_dfsspmdm_branch.restype = ctypes.c_int
        L.orc_dfsspmdm_plan.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5
        L.orc_dfsspmdm_plan.restype = ctypes.c_int
        L.orc_dfsspmdm_code_size.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4
        L.orc_dfsspmdm_code_size.restype = ctypes.c_long

    # -- spmdm ---------------------------------------------------------------------------
    def geometry(self, M, N, K, max_threads=1, bn=48):
        vec = (ctypes.c_int * 9)()
        scratch = self.lib.orc_spmdm_geometry(M, N, K, max_threads, bn, vec)
        return _geom_from(vec, scratch)

    def slices(self, g, a, transa="N", simd_w=None):
        a = np.ascontiguousarray(a)
        dtype = 0 if a.dtype == np.float32 else 1
        assert a.dtype in (np.float32, np.uint16)
        simd_w = simd_width_for_bn(g.bn) if simd_w is None else simd_w
        ro, co, va = empty_slices(g)
        self.lib.orc_spmdm_slices(g.vec, dtype, transa.encode(), _ptr(a), simd_w, _ptr(ro), _ptr(co), _ptr(va))
        return ro, co, va

    def compute(self, g, slices, b, c, transb="N", transc="N", beta=0.0, simd_w=None, tail_fma=1, fix_q17=False):
        """in-place on c (float32).  For bf16 pass beta as the float the reference derives
        from the raw bits: float(int(bits)).  fix_q17: scale ALL rows by beta in the AVX-512
        instantiation's transc='T' staging (the reference leaves rows 8..15 of each 16 x 16
        block unscaled, quirk Q17; mirrored by default because this is the restatement)."""
        b = np.ascontiguousarray(b)
        dtype = 0 if b.dtype == np.float32 else 1
        assert c.dtype == np.float32 and c.flags.c_contiguous
        simd_w = simd_width_for_bn(g.bn) if simd_w is None else simd_w
        ro, co, va = slices
        self.lib.orc_spmdm_compute(g.vec, dtype, transb.encode(), transc.encode(), float(beta),
                                   _ptr(ro), _ptr(co), _ptr(va), _ptr(b), _ptr(c), simd_w, int(tail_fma) | (2 if fix_q17 else 0))
        return c

    # -- fsspmdm -------------------------------------------------------------------------
    def dfsspmdm_plan(self, a, lda=None):
        a = np.ascontiguousarray(a, np.float64)
        M, K = a.shape
        lda = K if lda is None else lda
        rowptr = np.zeros(M + 1, np.uint32)
        colidx = np.zeros(M * K + 1, np.uint32)
        val = np.zeros(M * K + 1, np.float64)
        nu = ctypes.c_int(0)
        nnz = self.lib.orc_dfsspmdm_plan(M, K, lda, _ptr(a), _ptr(rowptr), _ptr(colidx), _ptr(val), ctypes.byref(nu))
        return dict(nnz=nnz, n_unique=nu.value, rowptr=rowptr, colidx=colidx[:nnz], values=val[:nnz])

    def dfsspmdm_code_size(self, a, ldb, ldc, beta):
        p = self.dfsspmdm_plan(a)
        return self.lib.orc_dfsspmdm_code_size(a.shape[0], _ptr(p["rowptr"]), _ptr(np.ascontiguousarray(p["colidx"])),
                                               p["n_unique"], ldb, ldc, int(beta == 1.0))

    def dfsspmdm_branch(self, a, ldb, ldc, beta, host_avx512=True):
        a = np.ascontiguousarray(a, np.float64)
        return self.lib.orc_dfsspmdm_branch(a.shape[0], a.shape[1], a.shape[1], _ptr(a), ldb, ldc, float(beta), int(host_avx512))

    def dfsspmdm_execute(self, a, B, C, beta, branch, N=None, ldb=None, ldc=None):
        a = np.ascontiguousarray(a, np.float64)
        M, K = a.shape
        ldb = B.shape[1] if ldb is None else ldb
        ldc = C.shape[1] if ldc is None else ldc
        N = B.shape[1] if N is None else N
        assert B.dtype == np.float64 and C.dtype == np.float64
        self.lib.orc_dfsspmdm_execute(M, N, K, K, ldb, ldc, float(beta), _ptr(a), _ptr(B), _ptr(C), int(branch))
        return C

    def sfsspmdm_execute(self, a, B, C, beta, N=None, ldb=None, ldc=None):
        a = np.ascontiguousarray(a, np.float32)
        M, K = a.shape
        ldb = B.shape[1] if ldb is None else ldb
        ldc = C.shape[1] if ldc is None else ldc
        N = B.shape[1] if N is None else N
        assert B.dtype == np.float32 and C.dtype == np.float32
        self.lib.orc_sfsspmdm_execute(M, N, K, K, ldb, ldc, float(beta), _ptr(a), _ptr(B), _ptr(C))
        return C


    # -- CSR x dense SoA (section 8f-1) ------------------------------------------------
    def csr_soa_execute(self, rowptr, colidx, values, B, C, N, ldb=None, ldc=None, soa=None, beta=0.0):
        """in place on C.  B: [E][K][ldb][soa], C: [E][M][ldc][soa] (E may be absent)."""
        values = np.ascontiguousarray(values)
        dbl = 1 if values.dtype == np.float64 else 0
        assert B.dtype == values.dtype and C.dtype == values.dtype and B.flags.c_contiguous and C.flags.c_contiguous
        if B.ndim == 3:
            B = B[None]; C = C[None]
        E, K, ldb_, soa_ = B.shape
        M = C.shape[1]
        rowptr = np.ascontiguousarray(rowptr, np.uint32); colidx = np.ascontiguousarray(colidx, np.uint32)
        f = self.lib.orc_csr_soa_execute
        f.argtypes = [ctypes.c_int] * 7 + [ctypes.c_double] + [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3
        f.restype = None
        f(dbl, M, N, K, ldb_ if ldb is None else ldb, C.shape[2] if ldc is None else ldc, soa_ if soa is None else soa, float(beta),
          _ptr(rowptr), _ptr(colidx), _ptr(values), _ptr(B), _ptr(C), E, K * ldb_ * soa_, M * C.shape[2] * soa_)
        return C


    def csr_soa_bsparse_execute(self, rowptr, colidx, values, A, C, N, beta=0.0):
        """B sparse (CSR over K rows): in place on C.  A: [E][M][lda][soa], C: [E][M][ldc][soa]."""
        values = np.ascontiguousarray(values)
        dbl = 1 if values.dtype == np.float64 else 0
        assert A.dtype == values.dtype and C.dtype == values.dtype and A.flags.c_contiguous and C.flags.c_contiguous
        if A.ndim == 3:
            A = A[None]; C = C[None]
        E, M, lda, soa = A.shape
        K = len(rowptr) - 1
        rowptr = np.ascontiguousarray(rowptr, np.uint32); colidx = np.ascontiguousarray(colidx, np.uint32)
        f = self.lib.orc_csr_soa_bsparse_execute
        f.argtypes = [ctypes.c_int] * 7 + [ctypes.c_double] + [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3
        f.restype = None
        f(dbl, M, N, K, lda, C.shape[2], soa, float(beta), _ptr(rowptr), _ptr(colidx), _ptr(values), _ptr(A), _ptr(C), E, M * lda * soa, M * C.shape[2] * soa)
        return C


    def csc_soa_execute(self, colptr, rowidx, values, A, C, N, beta=0.0, K=None):
        """B sparse in CSC (colptr over N columns): in place on C.  A: [E][M][lda][soa], C: [E][M][ldc][soa]; K = lda unless given."""
        values = np.ascontiguousarray(values)
        dbl = 1 if values.dtype == np.float64 else 0
        assert A.dtype == values.dtype and C.dtype == values.dtype and A.flags.c_contiguous and C.flags.c_contiguous
        if A.ndim == 3:
            A = A[None]; C = C[None]
        E, M, lda, soa = A.shape
        colptr = np.ascontiguousarray(colptr, np.uint32); rowidx = np.ascontiguousarray(rowidx, np.uint32)
        assert len(colptr) == N + 1
        f = self.lib.orc_csc_soa_execute
        f.argtypes = [ctypes.c_int] * 7 + [ctypes.c_double] + [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3
        f.restype = None
        f(dbl, M, N, lda if K is None else K, lda, C.shape[2], soa, float(beta), _ptr(colptr), _ptr(rowidx), _ptr(values), _ptr(A), _ptr(C), E, M * lda * soa, M * C.shape[2] * soa)
        return C


class Ref:
    """The compiled reference.  ``Ref.available()`` is False where oracle/_ref is absent."""

    @staticmethod
    def path(flavor="avx2"):
        return os.path.join(REF_DIR, "libref_driver%s.so" % ("" if flavor == "avx2" else "_" + flavor))

    @classmethod
    def available(cls, flavor="avx2"):
        return os.path.exists(cls.path(flavor))

    def __init__(self, flavor="avx2"):
        self.lib = L = ctypes.CDLL(self.path(flavor))
        L.refdrv_spmdm_run.restype = ctypes.c_int
        L.refdrv_spmdm_run.argtypes = ([ctypes.c_int] * 6 + [ctypes.c_char] * 3 + [ctypes.c_void_p] * 4 +
                                       [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_void_p] * 4)
        L.refdrv_fsspmdm_run.restype = ctypes.c_int
        L.refdrv_fsspmdm_run.argtypes = ([ctypes.c_int] * 6 + [ctypes.c_double] + [ctypes.c_void_p] * 3 +
                                         [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2)
        L.refdrv_spmdm_geometry.restype = ctypes.c_int
        L.refdrv_max_threads.restype = ctypes.c_int

    def max_threads(self):
        return self.lib.refdrv_max_threads()

    def csr_soa(self, rowptr, colidx, values, B, C, N, beta=0.0):
        """libxsmm_create_xcsr_soa + one kernel call per element, in place on C.  B: [E][K][ldb][soa], C: [E][M][ldc][soa].
        Returns the SoA width the reference's generator uses on this host (the arrays' last dimension must equal it)."""
        values = np.ascontiguousarray(values)
        dbl = 1 if values.dtype == np.float64 else 0
        assert B.dtype == values.dtype and C.dtype == values.dtype and B.flags.c_contiguous and C.flags.c_contiguous
        if B.ndim == 3:
            B = B[None]; C = C[None]
        E, K, ldb, soa = B.shape
        M, ldc = C.shape[1], C.shape[2]
        rowptr = np.ascontiguousarray(rowptr, np.uint32); colidx = np.ascontiguousarray(colidx, np.uint32)
        used = ctypes.c_int(0)
        f = self.lib.refdrv_csr_soa_run
        f.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double] + [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3 + [ctypes.c_void_p]
        f.restype = ctypes.c_int
        rc = f(dbl, M, N, K, ldb, ldc, float(beta), _ptr(rowptr), _ptr(colidx), _ptr(values), _ptr(B), _ptr(C), E, K * ldb * soa, M * ldc * soa,
               ctypes.cast(ctypes.byref(used), ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError("reference csr_soa kernel could not be generated (rc=%d)" % rc)
        if used.value != soa:
            raise ValueError("SoA width of the arrays is %d, the reference's generator uses %d on this host" % (soa, used.value))
        return used.value

    def csr_soa_bsparse(self, rowptr, colidx, values, A, C, N, beta=0.0):
        """libxsmm_create_xcsr_soa with B sparse (descriptor lda > 0, ldb = 0): A [E][M][lda][soa] dense, in place on C [E][M][ldc][soa]."""
        values = np.ascontiguousarray(values)
        dbl = 1 if values.dtype == np.float64 else 0
        assert A.dtype == values.dtype and C.dtype == values.dtype and A.flags.c_contiguous and C.flags.c_contiguous
        if A.ndim == 3:
            A = A[None]; C = C[None]
        E, M, lda, soa = A.shape
        K, ldc = len(rowptr) - 1, C.shape[2]
        rowptr = np.ascontiguousarray(rowptr, np.uint32); colidx = np.ascontiguousarray(colidx, np.uint32)
        used = ctypes.c_int(0)
        f = self.lib.refdrv_csr_soa_run_ex
        f.argtypes = [ctypes.c_int] * 7 + [ctypes.c_double] + [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3 + [ctypes.c_void_p]
        f.restype = ctypes.c_int
        rc = f(dbl, M, N, K, lda, 0, ldc, float(beta), _ptr(rowptr), _ptr(colidx), _ptr(values), _ptr(A), _ptr(C), E, M * lda * soa, M * ldc * soa,
               ctypes.cast(ctypes.byref(used), ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError("reference csr_soa (B sparse) kernel could not be generated (rc=%d)" % rc)
        if used.value != soa:
            raise ValueError("SoA width of the arrays is %d, the reference's generator uses %d on this host" % (soa, used.value))
        return used.value

    def csc_soa(self, colptr, rowidx, values, A, C, N, beta=0.0):
        """libxsmm_create_xcsc_soa (B sparse, CSC): A [E][M][lda][soa] dense (K = lda), in place on C [E][M][ldc][soa]."""
        values = np.ascontiguousarray(values)
        dbl = 1 if values.dtype == np.float64 else 0
        assert A.dtype == values.dtype and C.dtype == values.dtype and A.flags.c_contiguous and C.flags.c_contiguous
        if A.ndim == 3:
            A = A[None]; C = C[None]
        E, M, lda, soa = A.shape
        ldc = C.shape[2]
        colptr = np.ascontiguousarray(colptr, np.uint32); rowidx = np.ascontiguousarray(rowidx, np.uint32)
        used = ctypes.c_int(0)
        f = self.lib.refdrv_csc_soa_run
        f.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double] + [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3 + [ctypes.c_void_p]
        f.restype = ctypes.c_int
        rc = f(dbl, M, N, lda, lda, ldc, float(beta), _ptr(colptr), _ptr(rowidx), _ptr(values), _ptr(A), _ptr(C), E, M * lda * soa, M * ldc * soa,
               ctypes.cast(ctypes.byref(used), ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError("reference csc_soa kernel could not be generated (rc=%d)" % rc)
        if used.value != soa:
            raise ValueError("SoA width of the arrays is %d, the reference's generator uses %d on this host" % (soa, used.value))
        return used.value

    def csr_soa_bench(self, rowptr, colidx, values, B, C, N, beta=0.0, threads=0, reps=1):
        """the reference's SoA kernel over all elements with OpenMP, timed; returns seconds per repetition."""
        values = np.ascontiguousarray(values)
        dbl = 1 if values.dtype == np.float64 else 0
        E, K, ldb, soa = B.shape
        M, ldc = C.shape[1], C.shape[2]
        rowptr = np.ascontiguousarray(rowptr, np.uint32); colidx = np.ascontiguousarray(colidx, np.uint32)
        times = np.zeros(reps, np.float64)
        f = self.lib.refdrv_csr_soa_bench
        f.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double] + [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        f.restype = ctypes.c_int
        rc = f(dbl, M, N, K, ldb, ldc, float(beta), _ptr(rowptr), _ptr(colidx), _ptr(values), _ptr(B), _ptr(C), E, K * ldb * soa, M * ldc * soa,
               threads, reps, _ptr(times))
        if rc != 0:
            raise RuntimeError("reference csr_soa kernel could not be generated (rc=%d)" % rc)
        return times

    def csr_soa_bsparse_bench(self, rowptr, colidx, values, A, C, N, beta=0.0, threads=0, reps=1):
        """the reference's B-sparse SoA kernel over all elements with OpenMP, timed; returns seconds per repetition."""
        values = np.ascontiguousarray(values)
        dbl = 1 if values.dtype == np.float64 else 0
        E, M, lda, soa = A.shape
        K, ldc = len(rowptr) - 1, C.shape[2]
        rowptr = np.ascontiguousarray(rowptr, np.uint32); colidx = np.ascontiguousarray(colidx, np.uint32)
        times = np.zeros(reps, np.float64)
        f = self.lib.refdrv_csr_soa_bench_ex
        f.argtypes = [ctypes.c_int] * 7 + [ctypes.c_double] + [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        f.restype = ctypes.c_int
        rc = f(dbl, M, N, K, lda, 0, ldc, float(beta), _ptr(rowptr), _ptr(colidx), _ptr(values), _ptr(A), _ptr(C), E, M * lda * soa, M * ldc * soa,
               threads, reps, _ptr(times))
        if rc != 0:
            raise RuntimeError("reference csr_soa (B sparse) kernel could not be generated (rc=%d)" % rc)
        return times

    def mm_rm(self, a, B, C, beta=1.0, nblock=16, lda=None):
        """libxsmm_[sd]mmdispatch applied to a row-major panel like samples/pyfr/pyfr_gemm_rm.c:98-122; in place on C."""
        a = np.ascontiguousarray(a)
        dbl = 1 if a.dtype == np.float64 else 0
        assert B.dtype == a.dtype and C.dtype == a.dtype and B.flags.c_contiguous and C.flags.c_contiguous and B.shape[1] == C.shape[1]
        M, K = a.shape
        f = self.lib.refdrv_mm_rm_run
        f.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_int] + [ctypes.c_void_p] * 3
        f.restype = ctypes.c_int
        rc = f(dbl, M, B.shape[1], K, K if lda is None else lda, B.shape[1], float(beta), nblock, _ptr(a), _ptr(B), _ptr(C))
        if rc != 0:
            raise RuntimeError("reference mmdispatch failed rc=%d" % rc)
        return C

    @staticmethod
    def soa_width(dtype):
        """SoA width of the reference's generator on this host (generator_spgemm_csr_asparse_soa.c:126-156)."""
        avx512 = "avx512f" in open("/proc/cpuinfo").read()
        return (8 if avx512 else 4) if np.dtype(dtype) == np.float64 else (16 if avx512 else 8)

    def geometry(self, M, N, K, max_threads=1):
        vec = (ctypes.c_int * 9)()
        scratch = self.lib.refdrv_spmdm_geometry(M, N, K, max_threads, vec)
        return _geom_from(vec, scratch)

    def spmdm(self, a, b, c, M, N, K, transa="N", transb="N", transc="N", beta=0.0, threads=1,
              max_threads=None, reps=1, dump=True):
        """Runs slice+compute ``reps`` times in place on c.  beta: float for fp32 inputs, raw
        uint16 bits for bf16 inputs.  Returns (geometry, slices-or-None, times[reps,3])."""
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        dtype = 0 if a.dtype == np.float32 else 1
        assert a.dtype == b.dtype and c.dtype == np.float32 and c.flags.c_contiguous
        beta_arr = np.array([beta], np.float32 if dtype == 0 else np.uint16)
        max_threads = threads if max_threads is None else max_threads
        g = self.geometry(M, N, K, max_threads)
        sl = empty_slices(g) if dump else (None, None, None)
        times = np.zeros((reps, 3), np.float64)
        vec = (ctypes.c_int * 9)()
        rc = self.lib.refdrv_spmdm_run(dtype, M, N, K, threads, max_threads, transa.encode(), transb.encode(),
                                       transc.encode(), _ptr(a), _ptr(b), _ptr(beta_arr), _ptr(c), reps,
                                       _ptr(times), ctypes.cast(vec, ctypes.c_void_p), _ptr(sl[0]), _ptr(sl[1]), _ptr(sl[2]))
        if rc != 0:
            raise RuntimeError("reference spmdm failed rc=%d" % rc)
        return g, (sl if dump else None), times

    def fsspmdm(self, a, B, C, beta, N=None, ld=None, panel=64, threads=1, reps=1):
        """In place on C.  Returns (sparse_branch_taken, chunk, times[reps])."""
        dbl = 1 if a.dtype == np.float64 else 0
        a = np.ascontiguousarray(a)
        M, K = a.shape
        ld = B.shape[1] if ld is None else ld
        N = B.shape[1] if N is None else N
        assert B.dtype == a.dtype and C.dtype == a.dtype
        info = (ctypes.c_int * 2)()
        times = np.zeros(reps, np.float64)
        rc = self.lib.refdrv_fsspmdm_run(dbl, M, N, K, K, ld, float(beta), _ptr(a), _ptr(B), _ptr(C), panel, threads,
                                         reps, _ptr(times), ctypes.cast(info, ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError("reference fsspmdm failed rc=%d" % rc)
        return bool(info[0]), int(info[1]), times
