"""Developer tool: pinned H2D / D2H bandwidth of this box, alone and concurrently (through the C ABI)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
xs = importlib.import_module("libxsmm-1_b200")
L = xs.load()
n = 64 << 20
h1 = xs.HostBuffer((n,), np.uint8); h2 = xs.HostBuffer((n,), np.uint8)
d1 = xs.DeviceBuffer(n); d2 = xs.DeviceBuffer(n)
s1, s2 = xs.Stream(), xs.Stream()
def t(f, reps=10):
    f(); s1.synchronize(); s2.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    s1.synchronize(); s2.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: L.libxsmm_b200_memcpy_h2d_async(d1.ptr, h1.ptr, n, s1.ptr))
b = t(lambda: L.libxsmm_b200_memcpy_d2h_async(h2.ptr, d2.ptr, n, s2.ptr))
def both():
    L.libxsmm_b200_memcpy_h2d_async(d1.ptr, h1.ptr, n, s1.ptr); L.libxsmm_b200_memcpy_d2h_async(h2.ptr, d2.ptr, n, s2.ptr)
c = t(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  concurrent: %.2f ms for 64 MiB each way (%.1f GB/s per direction)" % (n / a / 1e9, n / b / 1e9, c * 1e3, n / c / 1e9))
