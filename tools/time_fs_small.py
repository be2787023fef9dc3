"""Developer tool: C3's operator on one GPU's share of the columns at 1/2/4/8 GPUs (2^20 .. 2^17), per block size."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
xs = importlib.import_module("libxsmm-1_b200")
for name in ("c3", "c5"):
    for shift in (0, 1, 2, 3):
        wl = dict(bench.WORKLOADS[name]); wl["N"] = wl["N"] >> shift
        gen = bench.run_fs_gpu(xs, wl, 20, 5, 1, want_e2e=False)
        assert next(gen) == "ready"
        r = next(gen)
        for _ in gen: pass
        print("%s N=2^%d block=%s: %.1f us  %.3f of HBM peak" % (name, wl["N"].bit_length() - 1, os.environ.get("LIBXSMM_B200_FS_BLOCK", "auto"), r["kernel_ms"] * 1e3, r["kernel_bytes"] / r["kernel_ms"] / 1e6 / 6554.2), flush=True)
