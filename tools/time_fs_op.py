"""Developer tool: times one of the 150 PyFR operators (tests/golden/operators_pyfr.npz) as dfsspmdm with N = 2^20 columns.
python tools/time_fs_op.py p5/tet/m0 [reps]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
xs = importlib.import_module("libxsmm-1_b200")
name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
d = np.load(os.path.join(ROOT, "tests", "golden", "operators_pyfr.npz"), allow_pickle=True)
i = [str(n) for n in d["names"]].index(name)
M, K = (int(x) for x in d["shapes"][i])
lo, hi = int(d["offsets"][i]), int(d["offsets"][i + 1])
a = np.zeros((M, K), np.float64)
a[d["rows"][lo:hi], d["cols"][lo:hi]] = d["vals"][lo:hi]
bench.fs_operator = lambda xs_, wl: a
for beta in (0.0, 1.0):
    wl = dict(kind="fsspmdm", M=M, K=K, dtype="f64", N=1 << 20, beta=beta, density=None, n_unique=None)
    gen = bench.run_fs_gpu(xs, wl, reps, 3, 1, want_e2e=False)
    assert next(gen) == "ready"
    r = next(gen)
    for _ in gen:
        pass
    print("%s %dx%d nnz=%d beta=%g: %.1f us  %.0f GB/s (%.3f of 6554)  %.0f GFLOP/s" % (name, M, K, hi - lo, beta, r["kernel_ms"] * 1e3, r["kernel_bytes"] / r["kernel_ms"] / 1e6,
          r["kernel_bytes"] / r["kernel_ms"] / 1e6 / 6554.2, r["flops"] / r["kernel_ms"] / 1e6))
