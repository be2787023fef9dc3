"""Developer aid: where does the structured-sparse kernel differ from A @ B?  python tools/debug_k4s.py M N K density beta"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LIBXSMM_B200_SPMDM_TC"] = "1"; os.environ["LIBXSMM_B200_TC16_SPARSE"] = "1"
xs = importlib.import_module("libxsmm-1_b200")
M, N, K = (int(x) for x in sys.argv[1:4]); density = float(sys.argv[4]); beta = float(sys.argv[5])
A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, density, dtype="bf16", seed=M + N + K)
p = xs.Spmdm(M, N, K, 1)
dA, dB, dC = (xs.DeviceBuffer.from_numpy(x) for x in (A, B, C0))
p.create_slices(dA, "N", True); p.compute(dB, dC, "N", "N", beta, True); xs.synchronize()
print(xs.last_compute_kernel())
C = dC.to_numpy(np.float32, C0.shape)
Af = xs.workloads.from_bf16_bits(A).astype(np.float64); Bf = xs.workloads.from_bf16_bits(B).astype(np.float64)
want = Af @ Bf + beta * C0
err = np.abs(C - want)
nz = (Af != 0).reshape(M, K // 4, 4).sum(2)
ovf_rows = np.nonzero((nz > 2).any(1))[0]
bad_rows = np.nonzero(err.max(1) > 1e-4 * np.abs(want).max())[0]
print("rows with overflow groups:", len(ovf_rows), "bad rows:", len(bad_rows), "bad and overflow:", len(np.intersect1d(ovf_rows, bad_rows)))
print("bad rows", bad_rows[:20], "max err", err.max(), "bad columns of first bad row:", np.nonzero(err[bad_rows[0]] > 1e-4)[0][:10] if len(bad_rows) else None)
if len(bad_rows):
    r = bad_rows[0]
    d = (C[r] - want[r])
    # which single product explains the difference?
    for k in np.nonzero(Af[r])[0]:
        if np.allclose(-Af[r, k] * Bf[k], d, atol=1e-3): print("row", r, "is missing k =", k, "group pattern", nz[r, k // 4], "positions", np.nonzero(Af[r, 4 * (k // 4):4 * (k // 4) + 4])[0])
