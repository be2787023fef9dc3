"""The CPU oracle (oracle/*.c) against the golden fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Runs anywhere: neither /root/reference nor a GPU is needed.
This is what pins the oracle; the GPU tests then compare the CUDA path with the oracle."""
import glob
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GEOM_KEYS = ("m", "n", "k", "bm", "bn", "bk", "mb", "nb", "kb")


def bits(x):
    return x.view({8: np.uint64, 4: np.uint32, 2: np.uint16}[x.dtype.itemsize])


def test_geometry_table(oracle):
    table = json.load(open(os.path.join(GOLDEN, "geometry.json")))
    assert len(table) >= 90
    for row in table:
        g = oracle.geometry(row["M"], row["N"], row["K"], row["T"], bn=row["bn"])
        for k in ("bm", "bn", "bk", "mb", "nb", "kb", "scratch"):
            assert g[k] == row[k], (row, dict(g))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "spmdm_*.npz"))), ids=os.path.basename)
def test_spmdm_golden(oracle, path):
    import pyoracle
    z = np.load(path)
    ta, tb, tc = [str(t) for t in z["trans"]]
    gv = [int(v) for v in z["geom"]]
    g = oracle.geometry(gv[0], gv[1], gv[2], int(z["threads"]), bn=gv[4])
    assert [g[k] for k in GEOM_KEYS] == gv
    sl = oracle.slices(g, z["A"], ta)
    cnt = pyoracle.slice_counts(g, z["rowidx"])
    for s in range(g.nslices):
        nrows = min(g.bm, g.m - (s % g.mb) * g.bm)
        np.testing.assert_array_equal(sl[0][s, :nrows + 1], z["rowidx"][s, :nrows + 1])
        np.testing.assert_array_equal(sl[1][s, :cnt[s]], z["colidx"][s, :cnt[s]])
        np.testing.assert_array_equal(bits(sl[2][s, :cnt[s]]), bits(z["values"][s, :cnt[s]]))
    C = z["C0"].copy()
    oracle.compute(g, sl, z["B"], C, tb, tc, float(z["beta"]))
    np.testing.assert_array_equal(bits(C), bits(z["C"]))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "fsspmdm_*.npz")) + glob.glob(os.path.join(GOLDEN, "pyfr_*.npz"))),
                         ids=os.path.basename)
def test_fsspmdm_golden(oracle, path):
    z = np.load(path)
    a, B, C0 = z["a"], z["B"], z["C0"]
    for beta in (0.0, 1.0):
        C = C0.copy()
        if a.dtype == np.float64:
            branch = oracle.dfsspmdm_branch(a, B.shape[1], C.shape[1], beta)
            assert bool(branch) == bool(z["sparse_branch"])
            oracle.dfsspmdm_execute(a, B, C, beta, branch)
        else:
            assert not bool(z["sparse_branch"])
            oracle.sfsspmdm_execute(a, B, C, beta)
        np.testing.assert_array_equal(bits(C), bits(z["C_beta%d" % int(beta)]))


def test_fsspmdm_branch_sweep(oracle, xs):
    """unique-value limit (31) and the 128 KiB code-size limit of the reference's generator; checked for
    the oracle AND for the product's host-side planner (no GPU needed)."""
    sweep = json.load(open(os.path.join(GOLDEN, "fsspmdm_branch.json")))
    w = xs.workloads
    for r in sweep:
        a = w.fsspmdm_operator(r["M"], r["K"], r["density"], r["n_unique"], np.float64, seed=r["seed"])
        assert bool(oracle.dfsspmdm_branch(a, r["ld"], r["ld"], r["beta"])) == r["sparse"], r
        p = xs.fsspmdm_plan(a, N=16, ldb=r["ld"], ldc=r["ld"], beta=r["beta"])
        assert p["sparse"] == r["sparse"] and p["chunk"] == r["chunk"], (r, p)
