"""GPU parity of K4s, the 2:4 structured-sparse tensor-core kernel of the bf16 spmdm compute step
(tcgen05.mma.sp, csrc/spmdm_compute_tc16s.cu) and of its CUDA-core overflow pass.

The kernel is picked for bf16 slices written by the wide slicing kernel (K1x: complete 128-column k-blocks, aligned rows)
when the host's density estimate is at most 2.2 %; LIBXSMM_B200_TC16_SPARSE=1 forces it at any density (every group of four
consecutive k with more than two nonzeros then goes through the overflow pass), =0 disables it.  Like the other tensor-core
kernels it keeps the contract of BASELINE.json for the bf16 path (1e-2 relative; 1e-5 is asserted), not the reference's
rounding sequence; the slices themselves stay bit-exact."""
import numpy as np
import pytest

from test_spmdm_gpu import gpu_spmdm, oracle_spmdm, valid_slices_equal

pytestmark = pytest.mark.gpu
K4S = "spmdm_compute_tc16s_kernel"


def rel(got, want):
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max() / max(float(np.abs(want).max()), 1e-30))


def force(monkeypatch, sparse="1"):
    monkeypatch.setenv("LIBXSMM_B200_SPMDM_TC", "1")
    monkeypatch.setenv("LIBXSMM_B200_TC16_SPARSE", sparse)


@pytest.mark.parametrize("M,N,K,density,beta,tc,threads", [
    (512, 512, 512, 0.01, 0, "N", 1),
    (512, 512, 512, 0.03, 1, "N", 1),
    (512, 512, 512, 0.10, 1, "N", 1),         # ~5 % of the groups overflow
    (512, 256, 256, 0.50, 0, "N", 1),         # most groups overflow: the CUDA-core pass does a large share of the work
    (640, 256, 128, 1.00, 0, "N", 1),         # full slices: every group overflows twice, u16 counter wraps, dense fallback of the workers
    (300, 208, 256, 0.05, 1, "N", 1),         # ragged M, partial column tile
    (2048, 264, 384, 0.02, 0, "N", 56),       # bm = 245: CTA tiles of 128 + 117 rows, pairs span row blocks
    (1536, 520, 384, 0.002, 0, "N", 1),       # nearly empty slices
    (33, 8, 128, 0.40, 0, "N", 1),            # one partial tile
    (256, 200, 384, 0.05, 1, "T", 1),         # C stored transposed
    (4096, 320, 256, 0.02, 0, "N", 1),        # more tiles than CTA pairs: several tiles per pair, single accumulator reused
])
def test_sparse_tensor_core_matches_oracle(gpu, oracle, monkeypatch, M, N, K, density, beta, tc, threads):
    force(monkeypatch)
    if density >= 1.0:
        rng = np.random.default_rng(5)
        A = gpu.workloads.to_bf16_bits((rng.random((M, K)) + 0.5).astype(np.float32))
        B = gpu.workloads.to_bf16_bits(rng.random((K, N)).astype(np.float32))
        C0 = rng.random((M, N)).astype(np.float32)
    else:
        A, B, C0 = gpu.workloads.spmdm_inputs(M, N, K, density, dtype="bf16", seed=M + N + K, transc=tc)
    g, sl, C = gpu_spmdm(gpu, A, B, C0, M, N, K, "N", "N", tc, beta, True, threads)
    assert gpu.last_compute_kernel() == K4S
    og, osl, OC = oracle_spmdm(oracle, g, A, B, C0, "N", "N", tc, float(beta))
    valid_slices_equal(og, sl, osl)
    assert rel(C, OC) <= 1e-5
    gpu.check()


def test_every_group_pattern(gpu, oracle, monkeypatch):
    """all 16 nonzero patterns of a group of four, at every group position of a 16-k span and in rows of both halves of a
    16-row metadata unit, with mixed signs: kept slots, nibbles and the overflow pass (patterns with 3 and 4 nonzeros)."""
    force(monkeypatch)
    M, N, K = 256, 256, 256
    rng = np.random.default_rng(11)
    A = np.zeros((M, K), np.float32)
    for r in range(M):
        for gidx in range(K // 4):
            pat = (r * 7 + gidx * 5 + (r >> 4)) % 16 if (r + gidx) % 3 == 0 else 0
            for p in range(4):
                if pat >> p & 1:
                    A[r, 4 * gidx + p] = rng.uniform(-1, 1)
    A[A == 0] = 0
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    w = gpu.workloads
    A16, B16 = w.to_bf16_bits(A), w.to_bf16_bits(B)
    g, sl, C = gpu_spmdm(gpu, A16, B16, C0, M, N, K, beta=1, bf16=True)
    assert gpu.last_compute_kernel() == K4S
    og, osl, OC = oracle_spmdm(oracle, g, A16, B16, C0, "N", "N", "N", 1.0)
    valid_slices_equal(og, sl, osl)
    # mixed signs: elementwise bound beside the normwise one (fp32 accumulation of exact products)
    want = w.from_bf16_bits(A16).astype(np.float64) @ w.from_bf16_bits(B16).astype(np.float64) + C0
    mag = np.abs(w.from_bf16_bits(A16)).astype(np.float64) @ np.abs(w.from_bf16_bits(B16)).astype(np.float64) + np.abs(C0)
    assert np.all(np.abs(C - want) <= 4e-6 * mag + 1e-30)
    assert rel(C, OC) <= 1e-5
    gpu.check()


def test_auto_dispatch_by_density(gpu, monkeypatch):
    """without switches: the first multiply of a handle has no density estimate (twin launch, K4p / CUDA cores); from the
    second on a 1.5 % matrix runs on the structured-sparse kernel, a 30 % one on K4p; same results within the contract."""
    monkeypatch.delenv("LIBXSMM_B200_SPMDM_TC", raising=False)
    monkeypatch.delenv("LIBXSMM_B200_TC16_SPARSE", raising=False)
    M = N = K = 1024
    xs = gpu
    for density, want_kernel in ((0.015, K4S), (0.30, "spmdm_compute_tc16p_kernel")):
        A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, density, dtype="bf16", seed=3)
        p = xs.Spmdm(M, N, K, 1)
        dA, dB, dC = (xs.DeviceBuffer.from_numpy(x) for x in (A, B, C0))
        outs = []
        for it in range(3):
            p.create_slices(dA, "N", True)
            xs.synchronize()                      # the density estimate is published when the slicing pass completes
            dC2 = xs.DeviceBuffer.from_numpy(C0)
            p.compute(dB, dC2, "N", "N", 0.0, True)
            xs.synchronize()
            outs.append((xs.last_compute_kernel(), dC2.to_numpy(np.float32, C0.shape)))
            dC2.free()
        assert outs[-1][0] == want_kernel, outs[-1][0]
        ref = xs.workloads.from_bf16_bits(A).astype(np.float64) @ xs.workloads.from_bf16_bits(B).astype(np.float64)
        for _, C in outs:
            assert rel(C, ref) <= 1e-5
        for d in (dA, dB, dC):
            d.free()
        p.destroy()
    gpu.check()


@pytest.mark.parametrize("density,beta", [(0.01, 0), (0.01, 1), (0.05, 0)])
def test_full_size_c2(gpu, monkeypatch, density, beta):
    """BASELINE C2 (bf16, 4096^3) on the structured-sparse kernel against the order-preserving CUDA-core kernels on the same
    device buffers (bit-exact against the oracle at the sizes the oracle can do), every output element."""
    M = N = K = 4096
    A, B, C0 = gpu.workloads.spmdm_inputs(M, N, K, density, dtype="bf16", seed=7)
    force(monkeypatch)
    _, _, C_sp = gpu_spmdm(gpu, A, B, C0, M, N, K, beta=beta, bf16=True)
    assert gpu.last_compute_kernel() == K4S
    monkeypatch.setenv("LIBXSMM_B200_SPMDM_TC", "0")
    _, _, C_cc = gpu_spmdm(gpu, A, B, C0, M, N, K, beta=beta, bf16=True)
    assert rel(C_sp, C_cc) <= 1e-5
    assert not np.array_equal(C_sp.view(np.uint32), C_cc.view(np.uint32))
    rows = np.arange(0, M, 37)
    want = gpu.workloads.from_bf16_bits(A[rows]).astype(np.float64) @ gpu.workloads.from_bf16_bits(B).astype(np.float64) + beta * C0[rows]
    assert rel(C_sp[rows], want) <= 1e-5
    gpu.check()


def test_handle_reuse_keeps_results_deterministic(gpu, monkeypatch):
    """the compressed tile and the metadata image are patched, never rebuilt: a second matrix through the same handle must
    not see anything of the first; two runs give the same bits (the overflow pass has one fixed order)."""
    force(monkeypatch)
    M, N, K = 1024, 512, 512
    xs = gpu
    p = xs.Spmdm(M, N, K, 1)
    res = []
    for seed, density in ((1, 0.08), (2, 0.01), (1, 0.08)):
        A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, density, dtype="bf16", seed=seed)
        dA, dB, dC = (xs.DeviceBuffer.from_numpy(x) for x in (A, B, C0))
        p.create_slices(dA, "N", True)
        p.compute(dB, dC, "N", "N", 0.0, True)
        xs.synchronize()
        assert xs.last_compute_kernel() == K4S
        C = dC.to_numpy(np.float32, C0.shape)
        ref = xs.workloads.from_bf16_bits(A).astype(np.float64) @ xs.workloads.from_bf16_bits(B).astype(np.float64)
        assert rel(C, ref) <= 1e-5
        res.append(C)
        for d in (dA, dB, dC):
            d.free()
    assert np.array_equal(res[0].view(np.uint32), res[2].view(np.uint32))
    p.destroy()
    gpu.check()


@pytest.mark.parametrize("shape,density,beta,tc", [((2048, 1024, 512), 0.01, 0, "N"), ((1200, 1000, 640), 0.02, 1, "N"), ((1024, 768, 384), 0.015, 0, "T")])
def test_exec_host_rectangles(gpu, oracle, monkeypatch, shape, density, beta, tc):
    """the whole-multiply host entry cuts C into row blocks x column panels: the structured-sparse kernel on sub-rectangles
    (first row block > 0, column origin > 0, ragged last panel), C boxes by TMA into a panel of a wider C."""
    force(monkeypatch)
    xs = gpu
    M, N, K = shape
    A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, density, dtype="bf16", seed=9, transc=tc)
    h, s = xs.libxsmm_spmdm_init(M, N, K, 1)
    try:
        import pyoracle
        g = pyoracle.Geometry(dict(m=h.m, n=h.n, k=h.k, bm=h.bm, bn=h.bn, bk=h.bk, mb=h.mb, nb=h.nb, kb=h.kb))
        og, osl, OC = oracle_spmdm(oracle, g, A, B, C0, "N", "N", tc, float(beta))
        for rep in range(2):
            C = C0.copy()
            xs.libxsmm_spmdm_exec_host(h, s, xs.LIBXSMM_SPMDM_DATATYPE_BFLOAT16, "N", "N", tc, A, B, beta, C)
            xs.check()
            assert xs.last_compute_kernel() == K4S
            assert rel(C, OC) <= 1e-5
    finally:
        xs.libxsmm_spmdm_destroy(h)


def test_changing_density_through_one_handle(gpu, monkeypatch):
    """a handle that sees matrices of very different density in turn (a cache keyed by shape, as TensorFlow's wrapper keeps):
    whenever the host's estimate is fresh or has just moved, the structured-sparse kernel is enqueued together with the dense one
    and the device picks by the slices' own counts -- a dense matrix never runs through the overflow path of K4s -- and every
    result is right; after two multiplies at about the same density K4s runs alone."""
    monkeypatch.delenv("LIBXSMM_B200_SPMDM_TC", raising=False)
    monkeypatch.delenv("LIBXSMM_B200_TC16_SPARSE", raising=False)
    M = N = K = 1024
    xs = gpu
    sets = {d: xs.workloads.spmdm_inputs(M, N, K, d, dtype="bf16", seed=int(d * 1000)) for d in (0.01, 0.30)}
    refs = {d: xs.workloads.from_bf16_bits(A).astype(np.float64) @ xs.workloads.from_bf16_bits(B).astype(np.float64) for d, (A, B, _) in sets.items()}
    p = xs.Spmdm(M, N, K, 1)
    names = []
    for d in (0.01, 0.30, 0.01, 0.30, 0.01, 0.01, 0.01, 0.01):
        A, B, C0 = sets[d]
        dA, dB, dC = (xs.DeviceBuffer.from_numpy(x) for x in (A, B, C0))
        p.create_slices(dA, "N", True)
        p.compute(dB, dC, "N", "N", 0.0, True)       # no synchronisation in between: the estimate is the previous pass's
        xs.synchronize()
        names.append(xs.last_compute_kernel())
        assert rel(dC.to_numpy(np.float32, C0.shape), refs[d]) <= 1e-5, (d, names)
        for b in (dA, dB, dC):
            b.free()
    assert names[-1] == K4S, names
    assert any(n.startswith("guarded:") for n in names), names
    p.destroy()
    gpu.check()
