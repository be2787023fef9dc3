"""Developer tool: dense 150x64 float operator, N = 2^22 columns: tensor-core kernel vs baked FMA kernel."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
xs = importlib.import_module("libxsmm-1_b200")
for tc in ((os.environ["LIBXSMM_B200_FSSPMDM_TC"],) if os.environ.get("LIBXSMM_B200_FSSPMDM_TC") in ("0", "1") else ("1", "0")):
    os.environ["LIBXSMM_B200_FSSPMDM_TC"] = tc
    for dens in ([float(x) for x in sys.argv[1:]] or [1.0, 0.5]):
        wl = dict(bench.WORKLOADS["c5"], density=dens, n_unique=None, N=1 << 22, beta=float(os.environ.get("FS_BETA", "0")))
        gen = bench.run_fs_gpu(xs, wl, 10, 3, 1, want_e2e=False)
        assert next(gen) == "ready"
        r = next(gen)
        for _ in gen: pass
        print("TC=%s density=%.1f: %.1f us  %.0f GB/s (%.2f of peak)  %.0f GFLOP/s" % (tc, dens, r["kernel_ms"] * 1e3, r["kernel_bytes"] / r["kernel_ms"] / 1e6, r["kernel_bytes"] / r["kernel_ms"] / 1e6 / 6554.2, r["flops"] / r["kernel_ms"] / 1e6))
