"""Developer tool: compute-step time of the structured-sparse tensor-core kernel (K4s) against the dense one (K4p) over density,
bf16 N/N/N, the measurement behind kSpMaxDensity in csrc/spmdm_compute_tc16s.cu.   python tools/crossover_sp.py [size]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
xs = importlib.import_module("libxsmm-1_b200")
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
os.environ["LIBXSMM_B200_SPMDM_TC"] = "1"
for dens in (0.005, 0.01, 0.02, 0.03, 0.05, 0.08, 0.12):
    wl = dict(kind="spmdm", M=size, N=size, K=size, density=dens, dtype="bf16", trans="NNN", beta=0, desc="x")
    out = []
    for sp in ("1", "0"):
        os.environ["LIBXSMM_B200_TC16_SPARSE"] = sp
        gen = bench.run_spmdm_gpu(xs, wl, 8, 3, want_e2e=False)
        next(gen); r = next(gen)
        for _ in gen: pass
        out.append("%s %.1f us (slice %.1f)" % (r["kernel_name"][14:30], r["parts"]["compute_ms"] * 1e3, r["parts"]["slice_ms"] * 1e3))
    print("bf16 %d^3 %.3f  " % (size, dens) + "   ".join(out), flush=True)
