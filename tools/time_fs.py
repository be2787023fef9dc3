"""Developer tool: times one fsspmdm workload (bench.py's c3 / c5) with CUDA events.  python tools/time_fs.py c5 [reps]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
xs = importlib.import_module("libxsmm-1_b200")
for beta in (0.0, 1.0):
    wl = dict(bench.WORKLOADS[name], beta=beta)
    gen = bench.run_fs_gpu(xs, wl, reps, 3, 1, want_e2e=False)
    assert next(gen) == "ready"
    r = next(gen)
    for _ in gen:
        pass
    print("%s beta=%g: %.1f us  %.0f GB/s (%.3f of 6554)  %.0f GFLOP/s  %s" % (name, beta, r["kernel_ms"] * 1e3, r["kernel_bytes"] / r["kernel_ms"] / 1e6,
          r["kernel_bytes"] / r["kernel_ms"] / 1e6 / 6554.2, r["flops"] / r["kernel_ms"] / 1e6, r["geo"]))
