"""Dense SMM dispatch used on row-major panels with very many columns: reference
libxsmm_[sd]mmdispatch driven like samples/pyfr/pyfr_gemm_rm.c:98-122.  CPU: the oracle's dense-branch restatement equals
the compiled reference's dispatched kernels bit for bit.  GPU: libxsmm_b200_[sd]mmdispatch + execute against the oracle."""
import numpy as np
import pytest


def bits(x):
    return x.view(np.uint64 if x.dtype == np.float64 else np.uint32)


def operator(rng, M, K, dt, density=1.0):
    a = rng.uniform(-1, 1, (M, K)).astype(dt)
    if density < 1.0:
        a[rng.random((M, K)) >= density] = 0
    a[M // 2, :] = 0          # an all-zero operator row: the SMM kernel still writes it (beta C or 0)
    return a


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_oracle_matches_dispatched_reference_kernels(oracle, ref, dt):
    rng = np.random.default_rng(2)
    for (M, K, N, beta, nblock) in ((150, 64, 96, 0.0, 16), (105, 60, 48, 1.0, 48), (20, 130, 64, 1.0, 32), (192, 96, 32, 0.0, 16)):
        a = operator(rng, M, K, dt, 0.6)
        B = rng.uniform(-1, 1, (K, N)).astype(dt); C0 = rng.uniform(-1, 1, (M, N)).astype(dt)
        C = C0.copy(); ref.mm_rm(a, B, C, beta, nblock=nblock)
        OC = C0.copy()
        if dt == np.float64:
            oracle.dfsspmdm_execute(a, B, OC, beta, 0)
        else:
            oracle.sfsspmdm_execute(a, B, OC, beta)
        assert np.array_equal(bits(C), bits(OC)), (M, K, N, beta)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_gpu_dispatch_matches_oracle(gpu, oracle, dt):
    rng = np.random.default_rng(4)
    xs = gpu
    for (M, K, N, ld, beta, density, nblock) in ((150, 64, 4096, 4096, 0.0, 0.2, 16), (105, 60, 1000, 1024, 1.0, 0.5, 48), (150, 125, 2080, 2080, 1.0, 0.1, 16),
                                                 (84, 252, 512, 512, 0.0, 0.96, 32)):
        a = operator(rng, M, K, dt, density)
        B = rng.uniform(-1, 1, (K, ld)).astype(dt); C0 = rng.uniform(-1, 1, (M, ld)).astype(dt)
        want = C0.copy()
        if dt == np.float64:
            oracle.dfsspmdm_execute(a, B, want, beta, 0, N=N, ldb=ld, ldc=ld)
        else:
            oracle.sfsspmdm_execute(a, B, want, beta, N=N, ldb=ld, ldc=ld)
        h = xs.MmDispatch(nblock, M, K, lda=ld, ldb=K, ldc=ld, beta=beta, dtype=dt)
        dB, dC = xs.DeviceBuffer.from_numpy(B), xs.DeviceBuffer.from_numpy(C0)
        h.execute(dB, a, dC, N)                       # operator from a HOST pointer
        xs.synchronize()
        C = dC.to_numpy(dt, C0.shape)
        assert h.kernel == "fs_baked", h.kernel
        assert np.array_equal(bits(C), bits(want)), (dt.__name__, M, K, N, beta)
        # the same handle with ANOTHER operator, this time from a device pointer: re-baked, still right
        a2 = operator(rng, M, K, dt, density)
        want2 = C0.copy()
        if dt == np.float64:
            oracle.dfsspmdm_execute(a2, B, want2, beta, 0, N=N, ldb=ld, ldc=ld)
        else:
            oracle.sfsspmdm_execute(a2, B, want2, beta, N=N, ldb=ld, ldc=ld)
        dA2 = xs.DeviceBuffer.from_numpy(a2)
        dC.upload(C0)
        h.execute(dB, dA2, dC, N)
        xs.synchronize()
        C = dC.to_numpy(dt, C0.shape)
        assert np.array_equal(bits(C), bits(want2))
        for d in (dB, dC, dA2):
            d.free()
        h.release()
    xs.check()


@pytest.mark.gpu
def test_gpu_dispatch_dense_float_goes_to_tensor_cores(gpu, oracle):
    """a dense 150 x 64 float operator is a real contraction: the handle holds the tcgen05 kernel (1e-5 contract)"""
    rng = np.random.default_rng(6)
    M, K, N = 150, 64, 8192
    a = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32); C0 = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    want = C0.copy(); oracle.sfsspmdm_execute(a, B, want, 1.0)
    h = gpu.MmDispatch(16, M, K, lda=N, ldb=K, ldc=N, dtype=np.float32)        # NULL alpha, beta: 1 and 1 like LIBXSMM_ALPHA / LIBXSMM_BETA
    dB, dC = gpu.DeviceBuffer.from_numpy(B), gpu.DeviceBuffer.from_numpy(C0)
    h.execute(dB, a, dC, N)
    gpu.synchronize()
    C = dC.to_numpy(np.float32, C0.shape)
    assert h.kernel == "fs_tc_kernel"
    assert float(np.abs(C.astype(np.float64) - want).max() / np.abs(want).max()) <= 1e-5
    dB.free(); dC.free(); h.release()
    with pytest.raises(ValueError):
        gpu.MmDispatch(16, M, K, lda=N, ldb=K, ldc=N, alpha=2.0)
    gpu.check()
