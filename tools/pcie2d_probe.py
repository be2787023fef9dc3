"""Developer tool: throughput of pinned 2-D copies (cudaMemcpy2DAsync) versus row width, each direction and both at once."""
import time
from cuda.bindings import runtime as rt

def ck(r):
    if isinstance(r, tuple):
        if r[0] != rt.cudaError_t.cudaSuccess: raise RuntimeError(str(r[0]))
        return r[1] if len(r) == 2 else r[1:]
    if r != rt.cudaError_t.cudaSuccess: raise RuntimeError(str(r))

n = 64 << 20
h1 = ck(rt.cudaHostAlloc(n, 0)); h2 = ck(rt.cudaHostAlloc(n, 0))
d1 = ck(rt.cudaMalloc(n)); d2 = ck(rt.cudaMalloc(n))
s1 = ck(rt.cudaStreamCreate()); s2 = ck(rt.cudaStreamCreate())
H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost

def timed(f, reps=8):
    f(); ck(rt.cudaDeviceSynchronize())
    t0 = time.perf_counter()
    for _ in range(reps): f()
    ck(rt.cudaDeviceSynchronize())
    return (time.perf_counter() - t0) / reps

pitch = 16384
for width in (512, 1024, 2048, 4096, 8192, 16384):
    rows = 4096
    nbytes = width * rows
    # a column panel of a pitch-wide matrix; 8 panels per call so that the transfer is 8 * nbytes
    def up():
        for i in range(8): ck(rt.cudaMemcpy2DAsync(d1 + (i * width) % pitch, pitch, h1 + (i * width) % pitch, pitch, width, rows, H2D, s1))
    def down():
        for i in range(8): ck(rt.cudaMemcpy2DAsync(h2 + (i * width) % pitch, pitch, d2 + (i * width) % pitch, pitch, width, rows, D2H, s2))
    def both():
        up(); down()
    a, b, c = timed(up), timed(down), timed(both)
    print("row %5d B x %d rows: H2D %.1f GB/s  D2H %.1f GB/s  both %.1f GB/s per direction" % (width, rows, 8 * nbytes / a / 1e9, 8 * nbytes / b / 1e9, 8 * nbytes / c / 1e9))
def up1(): ck(rt.cudaMemcpyAsync(d1, h1, n, H2D, s1))
def down1(): ck(rt.cudaMemcpyAsync(h2, d2, n, D2H, s2))
def both1(): up1(); down1()
a, b, c = timed(up1), timed(down1), timed(both1)
print("contiguous 64 MiB: H2D %.1f GB/s  D2H %.1f GB/s  both %.1f GB/s per direction" % (n / a / 1e9, n / b / 1e9, n / c / 1e9))
