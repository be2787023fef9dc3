"""Developer tool: whole spmdm steps (slices + multiply) back to back on one stream with NO events in between, which is what lets
programmatic dependent launch overlap the kernels' ramps.   python tools/time_step_plain.py c2 [reps]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
xs = importlib.import_module("libxsmm-1_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wl = bench.WORKLOADS[name]
bf16 = wl["dtype"] == "bf16"
A, B, C0 = bench.spmdm_host_inputs(xs, wl)
p = xs.Spmdm(wl["M"], wl["N"], wl["K"], 1)
ring = [(xs.DeviceBuffer.from_numpy(A), xs.DeviceBuffer.from_numpy(B), xs.DeviceBuffer(C0.nbytes)) for _ in range(3)]
st = xs.Stream()
ta, tb, tc = wl["trans"]
def step(i):
    a, b, c = ring[i % 3]
    p.create_slices(a, ta, bf16, st)
    p.compute(b, c, tb, tc, wl["beta"], bf16, st)
for i in range(5): step(i)
st.synchronize()
e0, e1 = xs.Event(), xs.Event()
e0.record(st)
for i in range(reps): step(i)
e1.record(st)
st.synchronize()
print("%s PDL=%s: %.1f us per step (%s)" % (name, os.environ.get("LIBXSMM_B200_PDL", "1"), e0.elapsed_ms(e1) / reps * 1e3, xs.last_compute_kernel()))
