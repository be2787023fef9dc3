"""Developer tool: bf16 spmdm through the tensor-core kernels (env selects) against the oracle."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle
xs = importlib.import_module("libxsmm-1_b200")
from test_spmdm_gpu import gpu_spmdm, oracle_spmdm
orc = pyoracle.Oracle()
for (M, N, K, d, beta) in [(128, 256, 128, 0.05, 0), (512, 512, 512, 0.01, 0), (300, 204, 256, 0.1, 1), (4096, 320, 256, 0.01, 0), (512, 512, 512, 0.5, 0), (1024, 1024, 1024, 0.01, 0)]:
    A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, d, dtype="bf16", seed=M + N)
    g, sl, C = gpu_spmdm(xs, A, B, C0, M, N, K, beta=beta, bf16=True)
    og, osl, OC = oracle_spmdm(orc, g, A, B, C0, "N", "N", "N", float(beta))
    err = np.abs(C.astype(np.float64) - OC.astype(np.float64)).max() / np.abs(OC).max()
    print("M=%d N=%d K=%d d=%.2f beta=%g: rel err %.3g  bit-equal %s  %s" % (M, N, K, d, beta, err, np.array_equal(C, OC), "OK" if err <= 1e-5 else "FAIL"), flush=True)
    xs.check()
