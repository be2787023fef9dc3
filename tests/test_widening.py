"""The rows SURVEY.md section 8(f) ranks next, at the same bar as the hot path: the MatrixMarket operator reader
(8f-2; reference src/generator_spgemm_csr_reader.c:46-169) and the fused caller step with a handle cache (8f-4;
reference documentation/tensorflow.md:241-250, samples/spmdm/spmdm.c:74-154)."""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def write_mtx(path, a, comment=True, order=None):
    rows, cols = np.nonzero(a)
    idx = np.arange(len(rows)) if order is None else order
    with open(path, "w") as f:
        if comment:
            f.write("%%MatrixMarket matrix coordinate real general\n% written by the test\n")
        f.write("%d %d %d\n" % (a.shape[0], a.shape[1], len(rows)))
        for i in idx:
            f.write("%d %d %.17g\n" % (rows[i] + 1, cols[i] + 1, a[rows[i], cols[i]]))


def test_mtx_reader_matches_the_harness_reader(xs, tmp_path):
    rng = np.random.default_rng(1)
    a = np.where(rng.random((37, 19)) < 0.2, rng.standard_normal((37, 19)), 0.0)
    a[5, :] = 0; a[36, :] = 0; a[0, :] = 0            # empty rows (reference reader: row_idx[i+1] = row_idx[i])
    p = str(tmp_path / "op.mtx")
    write_mtx(p, a)
    rp, ci, va, nr, nc = xs.csr_read_mtx(p)
    assert (nr, nc) == a.shape and rp[0] == 0 and rp[-1] == np.count_nonzero(a) and (np.diff(rp.astype(np.int64)) >= 0).all()
    dense = np.zeros_like(a)
    for r in range(nr):
        dense[r, ci[rp[r]:rp[r + 1]]] = va[rp[r]:rp[r + 1]]
    np.testing.assert_array_equal(dense, a)
    np.testing.assert_array_equal(xs.workloads.read_mtx(p), a)
    for r in range(nr):                                  # columns ascending inside a row, like the file
        assert (np.diff(ci[rp[r]:rp[r + 1]].astype(np.int64)) > 0).all()
    # entries in arbitrary order: bucketed by row (the reference reader requires row-sorted files)
    p2 = str(tmp_path / "shuffled.mtx")
    write_mtx(p2, a, comment=False, order=rng.permutation(np.count_nonzero(a)))
    rp2, ci2, va2, _, _ = xs.csr_read_mtx(p2)
    np.testing.assert_array_equal(rp2, rp)
    dense2 = np.zeros_like(a)
    for r in range(nr):
        dense2[r, ci2[rp2[r]:rp2[r + 1]]] = va2[rp2[r]:rp2[r + 1]]
    np.testing.assert_array_equal(dense2, a)


@pytest.mark.parametrize("text", ["", "% only a comment\n", "3 3\n1 1 1.0\n", "3 3 2\n1 1 1.0\n", "3 3 1\n4 1 1.0\n", "3 3 1\n1 0 1.0\n", "3 3 1\n1 1\n"])
def test_mtx_reader_rejects_malformed_files(xs, tmp_path, text):
    p = str(tmp_path / "bad.mtx")
    open(p, "w").write(text)
    with pytest.raises(ValueError):
        xs.csr_read_mtx(p)
    with pytest.raises(ValueError):
        xs.csr_read_mtx(str(tmp_path / "missing.mtx"))
    xs.clear_error()


@pytest.mark.gpu
def test_operator_from_mtx_matches_the_reference_outputs(gpu, tmp_path):
    """real PyFR operators (golden fixtures = outputs of the compiled reference): file -> create_mtx -> execute."""
    files = sorted(glob.glob(os.path.join(GOLDEN, "pyfr_*.npz")))
    assert files
    for f in files[:3]:
        z = np.load(f)
        a, B, C0 = z["a"], z["B"], z["C0"]
        p = str(tmp_path / (os.path.basename(f) + ".mtx"))
        write_mtx(p, a)
        for beta, key in ((0.0, "C_beta0"), (1.0, "C_beta1")):
            op = gpu.Fsspmdm.from_mtx(p, B.shape[1], beta=beta)
            try:
                assert (op.M, op.K) == a.shape and op.is_sparse == bool(z["sparse_branch"])
                dB, dC = gpu.DeviceBuffer.from_numpy(B), gpu.DeviceBuffer.from_numpy(C0)
                op.execute_stream(dB, dC)
                gpu.synchronize()
                C = dC.to_numpy(np.float64, C0.shape)
                dB.free(); dC.free()
            finally:
                op.destroy()
            np.testing.assert_array_equal(C.view(np.uint64), z[key].view(np.uint64))
    gpu.check()


@pytest.mark.gpu
def test_fused_sparse_matmul_with_handle_cache(gpu, oracle, monkeypatch):
    """one call = handle lookup + slices + compute; same bits as the two-phase stream entries (CUDA-core kernels)."""
    from test_spmdm_gpu import gpu_spmdm
    monkeypatch.setenv("LIBXSMM_B200_SPMDM_TC", "0")
    xs = gpu
    xs.sparse_matmul_cache_clear()
    st = xs.Stream()
    shapes = [(256, 192, 256, "f32", "N", "N", "N", 0.5), (300, 208, 130, "bf16", "N", "N", "N", 0), (256, 192, 256, "f32", "N", "T", "N", 0.0)]
    for rep in range(2):
        for (M, N, K, dtype, ta, tb, tc, beta) in shapes:
            A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, 0.1, dtype=dtype, seed=M + N + rep, transa=ta, transb=tb, transc=tc)
            dt = xs.LIBXSMM_SPMDM_DATATYPE_BFLOAT16 if dtype == "bf16" else xs.LIBXSMM_SPMDM_DATATYPE_F32
            dA, dB, dC = (xs.DeviceBuffer.from_numpy(x) for x in (A, B, C0))
            assert 0 == xs.libxsmm_b200_sparse_matmul(dt, ta, tb, tc, M, N, K, 1, dA, dB, beta, dC, st)
            st.synchronize()
            C = dC.to_numpy(np.float32, C0.shape)
            for d in (dA, dB, dC):
                d.free()
            _, _, want = gpu_spmdm(xs, A, B, C0, M, N, K, ta, tb, tc, beta, dtype == "bf16", 1)
            np.testing.assert_array_equal(C.view(np.uint32), want.view(np.uint32))
        assert xs.sparse_matmul_cache_entries() == 2          # (256,192,256) is shared by the first and third shape
    xs.sparse_matmul_cache_clear()
    assert xs.sparse_matmul_cache_entries() == 0
    xs.check()
