"""GPU parity of libxsmm_[sd]fsspmdm (through the C ABI) against the CPU oracle.

Bar: 1e-12 relative for fp64, 1e-5 for fp32 (BASELINE.json north_star).  Both branches keep the
reference's in-order fma chain, so bit equality is required as well.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_F64 = 1e-12
RTOL_F32 = 1e-5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_gpu(xs, a, B, C0, beta, N=None, ld=None, host=False):
    """returns (C, is_sparse, is_baked).  B is K x ld, C is M x ld; the first N columns are computed."""
    ld = B.shape[1] if ld is None else ld
    N = B.shape[1] if N is None else N
    op = xs.Fsspmdm(a, N, ldb=ld, ldc=ld, beta=beta)
    try:
        if host:
            C = C0.copy()
            op.execute(B, C)
        else:
            dB = xs.DeviceBuffer.from_numpy(B)
            dC = xs.DeviceBuffer.from_numpy(C0)
            op.execute_stream(dB, dC)
            xs.synchronize()
            C = dC.to_numpy(C0.dtype, C0.shape)
            dB.free(); dC.free()
        xs.check()
        return C, op.is_sparse, op.is_baked
    finally:
        op.destroy()


def oracle_run(oracle, a, B, C0, beta, N=None):
    C = C0.copy()
    if a.dtype == np.float64:
        branch = oracle.dfsspmdm_branch(a, B.shape[1], C.shape[1], beta)
        oracle.dfsspmdm_execute(a, B, C, beta, branch, N=N)
    else:
        branch = 0
        oracle.sfsspmdm_execute(a, B, C, beta, N=N)
    return C, bool(branch)


def same_bits(x, y):
    u = np.uint64 if x.dtype == np.float64 else np.uint32
    np.testing.assert_array_equal(x.view(u), y.view(u))


@pytest.mark.parametrize("n_unique", [8, 31, 32, None])
@pytest.mark.parametrize("beta", [0.0, 1.0])
@pytest.mark.parametrize("jit", ["1", "0"])
def test_dfsspmdm_matches_oracle(gpu, oracle, monkeypatch, n_unique, beta, jit):
    """150 x 64 at 30 %: <= 31 distinct values -> the reference's sparse_reg branch, otherwise its dense
    branch (quirk Q9).  jit=0 forces the generic (non-baked) kernel."""
    monkeypatch.setenv("LIBXSMM_B200_FSSPMDM_JIT", jit)
    a = gpu.workloads.fsspmdm_operator(150, 64, 0.30, n_unique, np.float64, seed=3)
    rng = np.random.default_rng(4)
    N = 4096 + 16
    B = rng.random((64, N)); C0 = rng.random((150, N))
    C, sparse, baked = run_gpu(gpu, a, B, C0, beta)
    OC, obranch = oracle_run(oracle, a, B, C0, beta)
    assert sparse == obranch == (n_unique is not None and n_unique <= 31)
    assert baked == (jit == "1")
    err = np.abs(C - OC).max() / np.abs(OC).max()
    assert err <= RTOL_F64
    same_bits(C, OC)


@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_sfsspmdm_matches_oracle(gpu, oracle, beta):
    a = gpu.workloads.fsspmdm_operator(150, 64, 0.30, 8, np.float32, seed=5)
    rng = np.random.default_rng(6)
    N = 8192
    B = rng.random((64, N), np.float32); C0 = rng.random((150, N), np.float32)
    C, sparse, baked = run_gpu(gpu, a, B, C0, beta)
    OC, _ = oracle_run(oracle, a, B, C0, beta)
    assert not sparse     # the reference never has a sparse kernel for float (quirk Q9)
    assert np.abs(C - OC).max() / np.abs(OC).max() <= RTOL_F32
    same_bits(C, OC)


@pytest.mark.parametrize("host", [False, True])
def test_panel_of_wider_matrix(gpu, oracle, host):
    """N < ldb = ldc: execute touches only its column panel (how PyFR calls it,
    samples/pyfr/pyfr_driver_asp_reg.c:268-272,297-308)."""
    a = gpu.workloads.fsspmdm_operator(96, 40, 0.2, 5, np.float64, seed=8)
    rng = np.random.default_rng(9)
    ld, N = 1000, 640
    B = rng.random((40, ld)); C0 = rng.random((96, ld))
    C, sparse, _ = run_gpu(gpu, a, B, C0, 1.0, N=N, ld=ld, host=host)
    OC, _ = oracle_run(oracle, a, B, C0, 1.0, N=N)
    same_bits(C, OC)
    same_bits(C[:, N:], C0[:, N:])


@pytest.mark.parametrize("n_unique", [4, None])
def test_empty_rows(gpu, oracle, n_unique):
    """sparse branch: rows without nonzeros are left untouched even for beta = 0 (quirk Q10);
    dense branch: they are zeroed."""
    a = gpu.workloads.fsspmdm_operator(60, 32, 0.3, n_unique, np.float64, seed=10)
    a[7, :] = 0; a[59, :] = 0; a[0, :] = 0
    rng = np.random.default_rng(11)
    B = rng.random((32, 256)); C0 = rng.random((60, 256)) + 1.0
    for host in (False, True):
        C, sparse, _ = run_gpu(gpu, a, B, C0, 0.0, host=host)
        OC, _ = oracle_run(oracle, a, B, C0, 0.0)
        same_bits(C, OC)
        if sparse:
            same_bits(C[7], C0[7])
        else:
            assert not C[7].any()


def test_nan_value_is_replaced_like_the_generator(gpu, oracle):
    """a NaN operator value "matches" the last entry of the generator's unique table and is replaced by
    it (reference src/generator_spgemm_csr_asparse_reg.c:125-150)."""
    a = gpu.workloads.fsspmdm_operator(20, 16, 0.5, 3, np.float64, seed=12)
    a[5, 3] = np.nan
    rng = np.random.default_rng(13)
    B = rng.random((16, 64)); C0 = rng.random((20, 64))
    C, sparse, _ = run_gpu(gpu, a, B, C0, 0.0)
    OC, _ = oracle_run(oracle, a, B, C0, 0.0)
    assert sparse
    same_bits(C, OC)
    assert not np.isnan(C).any()


def test_contract_violations_return_null(gpu):
    xs = gpu
    a = xs.workloads.fsspmdm_operator(8, 8, 0.5, 3, np.float64)
    for kw in (dict(N=24), dict(N=0), dict(N=16, beta=0.5), dict(N=16, alpha=2.0), dict(N=32, ldb=16), dict(N=32, ldc=16), dict(N=16, lda=4)):
        with pytest.raises(ValueError):
            xs.Fsspmdm(a, **kw)
    xs.clear_error()


def test_golden_pyfr_operators(gpu):
    """real PyFR operators shipped with the reference (samples/pyfr/mats), outputs of the compiled
    reference stored by tests/golden/make_golden.py."""
    files = sorted(glob.glob(os.path.join(GOLDEN, "pyfr_*.npz")))
    assert files, "golden fixtures missing"
    for f in files:
        z = np.load(f)
        a, B, C0 = z["a"], z["B"], z["C0"]
        for beta, key in ((0.0, "C_beta0"), (1.0, "C_beta1")):
            C, sparse, _ = run_gpu(gpu, a, B, C0, beta)
            assert sparse == bool(z["sparse_branch"])
            same_bits(C, z[key])


@pytest.mark.parametrize("il", ["0", "2", "8"])
def test_emitter_forms_give_the_same_bits(gpu, monkeypatch, il):
    """fma-heavy fp64 operators: every form of the baked kernel's emitter (plain; rows interleaved explicitly with the values
    in constant memory and three CTAs per SM asked for) keeps each row's own fma order, so every form reproduces the compiled
    reference's output bit for bit -- whichever one the timing at create would pick."""
    monkeypatch.setenv("LIBXSMM_B200_FSSPMDM_TUNE", "0")          # no timing: exactly the form named below is baked
    monkeypatch.setenv("LIBXSMM_B200_FSSPMDM_IL", il)
    monkeypatch.setenv("LIBXSMM_B200_FSSPMDM_CONST", "0" if il == "0" else "1")
    monkeypatch.setenv("LIBXSMM_B200_FSSPMDM_MINCTAS", "0" if il == "0" else "3")
    z = np.load(os.path.join(GOLDEN, "pyfr_p4_tet_m6.npz"))
    a, B, C0 = z["a"], z["B"], z["C0"]
    for beta, key in ((0.0, "C_beta0"), (1.0, "C_beta1")):
        C, sparse, baked = run_gpu(gpu, a, B, C0, beta)
        assert baked and not sparse
        same_bits(C, z[key])


@pytest.mark.parametrize("dtype,logn", [(np.float64, 20), (np.float32, 22)])
def test_full_size_properties(gpu, oracle, dtype, logn):
    """C3 (fp64, N = 2^20) and a 2^22-column slab of C5 (fp32): (1) a sample of columns equals the oracle
    bit for bit (columns are independent), (2) B = ones gives the operator's row sums in every column,
    (3) linearity in C0 for beta = 1."""
    xs = gpu
    a = xs.workloads.fsspmdm_operator(150, 64, 0.30, 8, dtype, seed=1)
    N = 1 << logn
    rng = np.random.default_rng(2)
    B = rng.random((64, N), np.float32).astype(dtype)
    C0 = rng.random((150, N), np.float32).astype(dtype)
    for beta in (0.0, 1.0):
        C, _, baked = run_gpu(xs, a, B, C0, beta)
        assert baked
        cols = np.unique(np.concatenate([np.arange(0, 64), np.arange(N - 64, N), rng.integers(0, N, 2048)]))
        Bs = np.ascontiguousarray(B[:, cols]); Cs = np.ascontiguousarray(C0[:, cols])
        pad = (-len(cols)) % 16
        if pad:
            Bs = np.ascontiguousarray(np.pad(Bs, ((0, 0), (0, pad)))); Cs = np.ascontiguousarray(np.pad(Cs, ((0, 0), (0, pad))))
        OC, _ = oracle_run(oracle, a, Bs, Cs, beta)
        same_bits(np.ascontiguousarray(C[:, cols]), np.ascontiguousarray(OC[:, :len(cols)]))
    ones = np.ones((64, N), dtype)
    C, _, _ = run_gpu(xs, a, ones, C0, 0.0)
    rs, _ = oracle_run(oracle, a, np.ones((64, 16), dtype), np.zeros((150, 16), dtype), 0.0)
    assert (C == rs[:, :1]).all()


@pytest.mark.parametrize("M,K,density", [(150, 64, 1.0), (150, 64, 0.6), (96, 40, 1.0), (192, 64, 0.9), (17, 5, 1.0)])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_sfsspmdm_tensor_core_branch(gpu, oracle, monkeypatch, M, K, density, beta):
    """dense float operators go to tcgen05 (3xTF32): the branch where the reference itself uses its dense SMM
    kernel and the apply really is a dense contraction.  Contract 1e-5 relative (not the fma-chain rounding)."""
    xs = gpu
    monkeypatch.setenv("LIBXSMM_B200_FSSPMDM_TC", "1")
    a = xs.workloads.fsspmdm_operator(M, K, density, None, np.float32, seed=21)
    rng = np.random.default_rng(22)
    ld, N = 4096 + 64, 4096 + 32      # panel of a wider matrix, last 128-column tile partial
    B = rng.random((K, ld), np.float32); C0 = rng.random((M, ld), np.float32)
    op = xs.Fsspmdm(a, N, ldb=ld, ldc=ld, beta=beta)
    try:
        assert op.is_tensor_core
        dB = xs.DeviceBuffer.from_numpy(B); dC = xs.DeviceBuffer.from_numpy(C0)
        op.execute_stream(dB, dC)
        xs.synchronize()
        C = dC.to_numpy(np.float32, C0.shape)
        dB.free(); dC.free()
    finally:
        op.destroy()
    OC = C0.copy()
    oracle.sfsspmdm_execute(a, B, OC, beta, N=N)
    err = np.abs(C.astype(np.float64) - OC.astype(np.float64)).max() / np.abs(OC).max()
    assert err <= RTOL_F32, err
    same_bits(np.ascontiguousarray(C[:, N:]), np.ascontiguousarray(C0[:, N:]))     # columns beyond N untouched
    xs.check()


def test_sfsspmdm_branch_choice(gpu, monkeypatch):
    """sparse float operators stay on the baked FMA kernel (HBM-bound), dense ones take the tensor cores."""
    monkeypatch.delenv("LIBXSMM_B200_FSSPMDM_TC", raising=False)
    xs = gpu
    sparse = xs.Fsspmdm(xs.workloads.fsspmdm_operator(150, 64, 0.3, 8, np.float32, seed=1), 1024)
    dense = xs.Fsspmdm(xs.workloads.fsspmdm_operator(150, 64, 1.0, None, np.float32, seed=1), 1024)
    wide = xs.Fsspmdm(xs.workloads.fsspmdm_operator(150, 128, 1.0, None, np.float32, seed=1), 1024)   # K > 64: not eligible
    try:
        assert not sparse.is_tensor_core and dense.is_tensor_core and not wide.is_tensor_core
    finally:
        sparse.destroy(); dense.destroy(); wide.destroy()
