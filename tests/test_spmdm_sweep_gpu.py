"""Seeded random sweep of libxsmm_spmdm over shapes, densities, orientations, beta and thread counts (= block
geometries), every dispatch mode, against the CPU oracle.  CUDA-core kernels: bit for bit.  Default dispatch and
forced tensor cores: the contract of BASELINE.json (1e-5 relative; bf16 products are exact, so the same bound holds)."""
import numpy as np
import pytest

from test_spmdm_gpu import gpu_spmdm, oracle_spmdm, valid_slices_equal

pytestmark = pytest.mark.gpu


def cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        M = int(rng.choice([17, 64, 130, 256, 300, 512, 777, 1100]))
        N = int(rng.choice([8, 48, 100, 203, 256, 384, 520]))
        K = int(rng.choice([5, 128, 129, 200, 256, 400]))
        density = float(rng.choice([0.003, 0.02, 0.1, 0.4, 0.9]))
        dtype = str(rng.choice(["f32", "bf16"]))
        ta, tb, tc = (str(rng.choice(["N", "T"])) for _ in range(3))
        if dtype == "bf16":
            beta = int(rng.choice([0, 1]))          # the bf16 entry reads beta as a raw integer (quirk Q2): 0 or 1 are the usable values
        else:
            beta = float(rng.choice([0.0, 1.0, 0.5]))
        threads = int(rng.choice([1, 1, 8, 56]))
        out.append((M, N, K, density, dtype, ta, tb, tc, beta, threads, i))
    return out


@pytest.mark.parametrize("M,N,K,density,dtype,ta,tb,tc,beta,threads,idx", cases(36, 20261018))
def test_random_sweep(gpu, oracle, monkeypatch, M, N, K, density, dtype, ta, tb, tc, beta, threads, idx):
    bf16 = dtype == "bf16"
    A, B, C0 = gpu.workloads.spmdm_inputs(M, N, K, density, dtype=dtype, seed=1000 + idx, transa=ta, transb=tb, transc=tc)
    monkeypatch.setenv("LIBXSMM_B200_SPMDM_TC", "0")
    g, sl, C = gpu_spmdm(gpu, A, B, C0, M, N, K, ta, tb, tc, beta, bf16, threads)
    og, osl, OC = oracle_spmdm(oracle, g, A, B, C0, ta, tb, tc, float(beta))
    valid_slices_equal(og, sl, osl)
    np.testing.assert_array_equal(C.view(np.uint32), OC.view(np.uint32))
    scale = max(float(np.abs(OC).max()), 1e-30)
    for mode in ("1", None):
        if mode is None:
            monkeypatch.delenv("LIBXSMM_B200_SPMDM_TC", raising=False)
        else:
            monkeypatch.setenv("LIBXSMM_B200_SPMDM_TC", mode)
        g2, sl2, C2 = gpu_spmdm(gpu, A, B, C0, M, N, K, ta, tb, tc, beta, bf16, threads)
        valid_slices_equal(og, sl2, osl)
        err = float(np.abs(C2.astype(np.float64) - OC.astype(np.float64)).max()) / scale
        assert err <= 1e-5, "mode %r: relative error %g (%s)" % (mode, err, gpu.last_compute_kernel())
    gpu.check()
