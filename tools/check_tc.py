"""Developer tool: spmdm with the tensor-core branch forced (LIBXSMM_B200_SPMDM_TC=1) against the oracle."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle
os.environ.setdefault("LIBXSMM_B200_SPMDM_TC", "1")
xs = importlib.import_module("libxsmm-1_b200")
from test_spmdm_gpu import gpu_spmdm, oracle_spmdm
orc = pyoracle.Oracle()
for (M, N, K, d, beta) in [(128, 128, 128, 0.5, 0.0), (256, 256, 256, 0.5, 0.0), (300, 203, 260, 0.3, 0.5), (1024, 512, 512, 0.5, 1.0), (2048, 2048, 2048, 0.1, 0.0)]:
    A, B, C0 = xs.workloads.spmdm_inputs(M, N, K, d, seed=M + N)
    g, sl, C = gpu_spmdm(xs, A, B, C0, M, N, K, beta=beta)
    if M <= 1024:
        og, osl, OC = oracle_spmdm(orc, g, A, B, C0, "N", "N", "N", beta)
    else:
        OC = (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32) + np.float32(beta) * C0
    err = np.abs(C.astype(np.float64) - OC.astype(np.float64)).max() / np.abs(OC).max()
    print("M=%d N=%d K=%d d=%.2f beta=%g: rel err %.3g  %s" % (M, N, K, d, beta, err, "OK" if err <= 1e-5 else "FAIL"), flush=True)
    xs.check()
