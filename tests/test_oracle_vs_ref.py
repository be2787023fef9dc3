"""The CPU oracle against the UNMODIFIED reference compiled by oracle/build_ref.sh -- bit for bit.
Needs oracle/_ref (built here from /root/reference; travels to the GPU box as a prebuilt .so).
Skipped where neither exists."""
import numpy as np
import pytest


def bits(x):
    return x.view({8: np.uint64, 4: np.uint32, 2: np.uint16}[x.dtype.itemsize])


SP = [
    (300, 200, 260, 0.10, "f32", "N", "N", "N", 0.0, 1),
    (300, 203, 260, 0.10, "f32", "N", "N", "N", 0.5, 3),
    (257, 190, 300, 0.50, "f32", "T", "N", "T", 1.0, 1),
    (257, 190, 300, 0.50, "f32", "N", "T", "N", 0.0, 2),
    (257, 77, 300, 0.30, "f32", "T", "T", "T", 0.25, 1),
    (1024, 96, 512, 0.15, "f32", "N", "N", "N", 0.0, 8),
    (512, 512, 512, 0.01, "bf16", "N", "N", "N", 0, 1),
    (300, 203, 256, 0.10, "bf16", "N", "N", "N", 1, 2),
    (256, 200, 384, 0.20, "bf16", "T", "N", "T", 0, 1),
    (256, 200, 384, 0.20, "bf16", "N", "T", "N", 0, 1),
]


@pytest.mark.parametrize("M,N,K,d,dt,ta,tb,tc,beta,T", SP)
def test_spmdm(oracle, ref, xs, M, N, K, d, dt, ta, tb, tc, beta, T):
    import pyoracle
    A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, d, dtype=dt, seed=M ^ N, transa=ta, transb=tb, transc=tc)
    C = C0.copy()
    g, sl, _ = ref.spmdm(A, B, C, M, N, K, ta, tb, tc, beta, threads=T)
    og = oracle.geometry(M, N, K, T, bn=g.bn)
    assert dict(og) == dict(g)
    # the product's host-side geometry is the same arithmetic
    pg = xs.spmdm_geometry(M, N, K, T, g.bn)
    assert pg == {k: g[k] for k in pg}
    osl = oracle.slices(og, A, ta)
    cnt = pyoracle.slice_counts(g, sl[0])
    for s in range(g.nslices):
        nrows = min(g.bm, g.m - (s % g.mb) * g.bm)
        np.testing.assert_array_equal(osl[0][s, :nrows + 1], sl[0][s, :nrows + 1])
        np.testing.assert_array_equal(osl[1][s, :cnt[s]], sl[1][s, :cnt[s]])
        np.testing.assert_array_equal(bits(osl[2][s, :cnt[s]]), bits(sl[2][s, :cnt[s]]))
    OC = C0.copy()
    oracle.compute(og, osl, B, OC, tb, tc, float(beta))
    np.testing.assert_array_equal(bits(OC), bits(C))


def test_spmdm_wrapped_counter(oracle, ref):
    """fully dense 512 x 128 slice: the u16 counter wraps (quirk Q4)."""
    M, N, K = 512, 48, 128
    A = (np.random.default_rng(2).random((M, K)) + 0.5).astype(np.float32)
    B = np.random.default_rng(3).random((K, N)).astype(np.float32)
    C = np.zeros((M, N), np.float32)
    g, sl, _ = ref.spmdm(A, B, C, M, N, K, beta=0.0)
    assert g.bm == 512 and sl[0][0, 512] == 0
    osl = oracle.slices(oracle.geometry(M, N, K, 1, bn=g.bn), A)
    np.testing.assert_array_equal(osl[0], sl[0])


@pytest.mark.parametrize("nu", [1, 8, 31, 32, None])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_dfsspmdm(oracle, ref, xs, nu, beta):
    a = xs.workloads.fsspmdm_operator(150, 64, 0.3, nu, np.float64, seed=3)
    a[11, :] = 0
    rng = np.random.default_rng(4)
    B = rng.random((64, 256)); C0 = rng.random((150, 256))
    C = C0.copy()
    sparse, chunk, _ = ref.fsspmdm(a, B, C, beta, panel=64)
    branch = oracle.dfsspmdm_branch(a, 256, 256, beta)
    assert bool(branch) == sparse
    OC = C0.copy()
    oracle.dfsspmdm_execute(a, B, OC, beta, branch)
    np.testing.assert_array_equal(bits(OC), bits(C))


def test_dfsspmdm_nan_value(oracle, ref, xs):
    a = xs.workloads.fsspmdm_operator(20, 16, 0.5, 3, np.float64, seed=12)
    a[5, 3] = np.nan
    rng = np.random.default_rng(13)
    B = rng.random((16, 64)); C0 = rng.random((20, 64)); C = C0.copy()
    sparse, _, _ = ref.fsspmdm(a, B, C, 0.0, panel=64)
    OC = C0.copy()
    oracle.dfsspmdm_execute(a, B, OC, 0.0, oracle.dfsspmdm_branch(a, 64, 64, 0.0))
    np.testing.assert_array_equal(bits(OC), bits(C))


@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_sfsspmdm(oracle, ref, xs, beta):
    a = xs.workloads.fsspmdm_operator(150, 64, 0.3, 8, np.float32, seed=5)
    rng = np.random.default_rng(6)
    B = rng.random((64, 256), np.float32); C0 = rng.random((150, 256), np.float32); C = C0.copy()
    sparse, chunk, _ = ref.fsspmdm(a, B, C, beta, panel=64)
    assert not sparse and chunk == 16
    OC = C0.copy()
    oracle.sfsspmdm_execute(a, B, OC, beta)
    np.testing.assert_array_equal(bits(OC), bits(C))


# ---- the reference's other two instantiations (src/libxsmm_spmdm.c:557-583) -----------------------------------------------
# AVX-512 (bn = 96, 16-wide vectors): `oracle/build_ref.sh avx512`.  Scalar (bn = 6): the AVX2 build run with
# LIBXSMM_TARGET=sse -- fp32 only, the scalar bf16 path of the reference is wrong.  They pin the
# restatement's mode boundaries (vector part / per-k-block partial sums / scalar tail) for simd_w = 16 and 1.
@pytest.fixture(scope="session")
def ref512():
    import os
    import pyoracle
    if "avx512f" not in open("/proc/cpuinfo").read():
        pytest.skip("host has no AVX-512")
    if not pyoracle.Ref.available("avx512") and os.path.isdir("/root/reference/src"):
        pyoracle.build_ref("avx512")
    if not pyoracle.Ref.available("avx512"):
        pytest.skip("oracle/_ref avx512 flavour not built (no /root/reference here)")
    return pyoracle.Ref("avx512")


@pytest.mark.parametrize("M,N,K,d,dt,ta,tb,tc,beta,T", SP + [(300, 330, 260, 0.2, "f32", "N", "N", "N", 0.5, 1), (512, 250, 256, 0.1, "bf16", "N", "N", "N", 0, 1)])
def test_spmdm_avx512_instantiation(oracle, ref512, xs, M, N, K, d, dt, ta, tb, tc, beta, T):
    import pyoracle
    A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, d, dtype=dt, seed=M ^ N, transa=ta, transb=tb, transc=tc)
    C = C0.copy()
    g, sl, _ = ref512.spmdm(A, B, C, M, N, K, ta, tb, tc, beta, threads=T)
    assert g.bn == 96
    og = oracle.geometry(M, N, K, T, bn=96)
    assert dict(og) == dict(g)
    assert xs.spmdm_geometry(M, N, K, T, 96) == {k: g[k] for k in ("m", "n", "k", "bm", "bn", "bk", "mb", "nb", "kb")}
    osl = oracle.slices(og, A, ta)
    cnt = pyoracle.slice_counts(g, sl[0])
    for s in range(g.nslices):
        nrows = min(g.bm, g.m - (s % g.mb) * g.bm)
        np.testing.assert_array_equal(osl[0][s, :nrows + 1], sl[0][s, :nrows + 1])
        np.testing.assert_array_equal(osl[1][s, :cnt[s]], sl[1][s, :cnt[s]])
        np.testing.assert_array_equal(bits(osl[2][s, :cnt[s]]), bits(sl[2][s, :cnt[s]]))
    OC = C0.copy()
    oracle.compute(og, osl, B, OC, tb, tc, float(beta))
    np.testing.assert_array_equal(bits(OC), bits(C))


SSE_WORKER = r'''
import os, sys
os.environ["LIBXSMM_TARGET"] = "sse"          # read once by the reference's init (src/libxsmm_main.c:618)
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import importlib, pyoracle
w = importlib.import_module("libxsmm-1_b200.workloads")
ref, orc = pyoracle.Ref(), pyoracle.Oracle()
for (M, N, K, d, ta, tb, tc, beta, T) in [(300, 200, 260, 0.1, "N", "N", "N", 0.0, 1), (300, 203, 260, 0.1, "N", "N", "N", 0.5, 3),
                                          (257, 190, 300, 0.5, "T", "N", "T", 1.0, 1), (257, 190, 300, 0.5, "N", "T", "N", 0.0, 2), (64, 7, 130, 0.3, "N", "N", "N", 0.25, 1)]:
    A, B, C0 = w.spmdm_inputs(M, N, K, d, dtype="f32", seed=M ^ N, transa=ta, transb=tb, transc=tc)
    A[0, 0] = np.nan                           # the scalar path KEEPS NaN (quirk Q3)
    C = C0.copy()
    g, sl, _ = ref.spmdm(A, B, C, M, N, K, ta, tb, tc, beta, threads=T)
    assert g.bn == 6, g
    og = orc.geometry(M, N, K, T, bn=6)
    assert dict(og) == dict(g), (og, g)
    osl = orc.slices(og, A, ta)
    cnt = pyoracle.slice_counts(g, sl[0])
    for s in range(g.nslices):
        nrows = min(g.bm, g.m - (s %% g.mb) * g.bm)
        assert np.array_equal(osl[0][s, :nrows + 1], sl[0][s, :nrows + 1])
        assert np.array_equal(osl[1][s, :cnt[s]], sl[1][s, :cnt[s]])
        assert np.array_equal(osl[2][s, :cnt[s]].view(np.uint32), sl[2][s, :cnt[s]].view(np.uint32))
    OC = C0.copy()
    orc.compute(og, osl, B, OC, tb, tc, float(beta))
    same = (OC.view(np.uint32) == C.view(np.uint32)) | (np.isnan(OC) & np.isnan(C))
    assert same.all(), "scalar instantiation: %%d elements differ" %% int((~same).sum())
print("sse ok")
'''


def test_spmdm_scalar_instantiation(oracle, ref, tmp_path):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "sse_worker.py"
    script.write_text(SSE_WORKER % {"root": root})
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "sse ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
