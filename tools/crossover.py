"""Developer tool: compute-step time of the CUDA-core and tensor-core kernels over density and orientation (2048^3),
the measurement behind tc_density_threshold() in csrc/common.cuh.   python tools/crossover.py [f32|bf16]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
xs = importlib.import_module("libxsmm-1_b200")
dtypes = sys.argv[1:] or ["f32", "bf16"]
for dtype in dtypes:
    for trans in ("NNN", "TNT", "NTN"):
        for dens in (0.005, 0.02, 0.04, 0.06, 0.08):
            wl = dict(kind="spmdm", M=2048, N=2048, K=2048, density=dens, dtype=dtype, trans=trans, beta=0 if dtype == "bf16" else 0.0, desc="x")
            out = []
            for tc in ("0", "1"):
                os.environ["LIBXSMM_B200_SPMDM_TC"] = tc
                gen = bench.run_spmdm_gpu(xs, wl, 8, 3, want_e2e=False)
                next(gen); r = next(gen)
                for _ in gen: pass
                out.append("TC=%s %.1f us (%s)" % (tc, r["parts"]["compute_ms"] * 1e3, r["kernel_name"][14:30]))
            print("%s %s %.3f  " % (dtype, trans, dens) + "   ".join(out), flush=True)
