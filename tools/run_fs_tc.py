"""Developer tool: one dense 150x64 float operator applied to 2^20 columns on the tensor-core kernel (for ncu)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
os.environ["LIBXSMM_B200_FSSPMDM_TC"] = "1"
xs = importlib.import_module("libxsmm-1_b200")
a = xs.workloads.fsspmdm_operator(150, 64, 1.0, None, np.float32, seed=1)
N = 1 << 20
op = xs.Fsspmdm(a, N)
dB = xs.DeviceBuffer(64 * N * 4); dC = xs.DeviceBuffer(150 * N * 4)
dB.fill(0); dC.fill(0)
for _ in range(4):
    op.execute_stream(dB, dC)
xs.synchronize()
print("ok", op.is_tensor_core)
