"""Developer tool: runs the C2 compute kernel back to back for ~2 s while sampling nvidia-smi clocks."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
xs = importlib.import_module("libxsmm-1_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[name]
bf16 = wl["dtype"] == "bf16"
A, B, C0 = bench.spmdm_host_inputs(xs, wl)
p = xs.Spmdm(wl["M"], wl["N"], wl["K"], 1)
ring = [(xs.DeviceBuffer.from_numpy(A), xs.DeviceBuffer.from_numpy(B), xs.DeviceBuffer(C0.nbytes)) for _ in range(3)]
st = xs.Stream()
p.create_slices(ring[0][0], "N", bf16, st)
st.synchronize()
s = bench.ClockSampler(0); s.start(); time.sleep(0.5)
for rep in range(4):
    e0, e1 = xs.Event(), xs.Event()
    n = 1000
    e0.record(st)
    for i in range(n):
        p.compute(ring[i % 3][1], ring[i % 3][2], "N", "N", wl["beta"], bf16, st)
    e1.record(st); st.synchronize()
    print("batch %d: %.1f us/launch" % (rep, e0.elapsed_ms(e1) / n * 1e3))
print(s.finish()); print(s.rows[:3], s.rows[-3:])
