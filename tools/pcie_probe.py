"""Developer tool: the host <-> device copy ceiling of this box: pinned H2D / D2H bandwidth alone and both directions at once,
for ONE rank or for N ranks copying concurrently (one process per GPU):

    python tools/pcie_probe.py                                                    # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

All ranks start every measurement together (barrier) and the slowest rank's time counts, so the printed aggregate is what N
concurrent end-to-end calls can move at best.  bench.py's e2e figures at N GPUs are to be read against it."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
xs = importlib.import_module("libxsmm-1_b200")
L = xs.load()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L.libxsmm_b200_set_device(local)
n = 64 << 20
h1 = xs.HostBuffer((n,), np.uint8); h2 = xs.HostBuffer((n,), np.uint8)
h1.array[...] = 1; h2.array[...] = 2
d1 = xs.DeviceBuffer(n); d2 = xs.DeviceBuffer(n)
s1, s2 = xs.Stream(), xs.Stream()

def barrier():
    if dist is not None:
        dist.barrier()

def t(f, reps=10):
    f(); s1.synchronize(); s2.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    s1.synchronize(); s2.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if dist is not None:
        import torch
        x = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        dt = float(x.item())
    return dt

a = t(lambda: L.libxsmm_b200_memcpy_h2d_async(d1.ptr, h1.ptr, n, s1.ptr))
b = t(lambda: L.libxsmm_b200_memcpy_d2h_async(h2.ptr, d2.ptr, n, s2.ptr))
def both():
    L.libxsmm_b200_memcpy_h2d_async(d1.ptr, h1.ptr, n, s1.ptr); L.libxsmm_b200_memcpy_d2h_async(h2.ptr, d2.ptr, n, s2.ptr)
c = t(both)
if 0 == rank:
    print(json.dumps({"ranks": world, "bytes_per_copy": n,
                      "h2d_gbs_per_rank": n / a / 1e9, "d2h_gbs_per_rank": n / b / 1e9, "both_gbs_per_direction_per_rank": n / c / 1e9,
                      "h2d_gbs_aggregate": world * n / a / 1e9, "d2h_gbs_aggregate": world * n / b / 1e9,
                      "both_gbs_aggregate_both_directions": 2 * world * n / c / 1e9,
                      "c2_e2e_floor_ms": 64 * 1.048576 / (n / c / 1e9)}))
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
