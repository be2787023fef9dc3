"""Column sharding on REAL devices: one process per GPU (torchrun, NCCL), rank r owns columns [n0, n0 + w) of B and C
(its own dense panel, ld = w) and a replica of the operator / of A; nothing is exchanged on the data path.  For
validation only, the C panels are all-gathered over NCCL (sharding.gather_columns) and rank 0 compares the assembled C
with the CPU oracle's C of the FULL problem, bit for bit (columns are independent, so the panel split must not change a
single bit).  Mirrors how the reference's driver walks column panels (samples/pyfr/pyfr_driver_asp_reg.c:297-308).

Needs >= 2 visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_sharding_nccl_gpu.py -m gpu`); skipped on a
one-GPU box.  The CPU-side logic of the same path is covered by tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import pyoracle
xs = importlib.import_module("libxsmm-1_b200")
sh = importlib.import_module("libxsmm-1_b200.sharding")
w = xs.workloads
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
xs.load(); xs.require_gpu()
assert 0 == xs.load().libxsmm_b200_set_device(local)
os.environ["LIBXSMM_B200_SPMDM_TC"] = "0"        # the order-preserving kernels: bit equality with the oracle is the bar
orc = pyoracle.Oracle()
dev = torch.device("cuda", local)

def panel(x, n0, wd):
    return torch.from_numpy(np.ascontiguousarray(x[:, n0:n0 + wd])).to(dev)

# ---- dfsspmdm / sfsspmdm: N columns in panels that are multiples of 16, uneven on purpose ---------------------------
for dtype, tdt, beta in ((np.float64, torch.float64, 1.0), (np.float64, torch.float64, 0.0), (np.float32, torch.float32, 1.0)):
    a = w.fsspmdm_operator(150, 64, 0.3, 8, dtype, seed=1)
    N = 16 * (1000 * world + 3)
    rng = np.random.default_rng(2)
    B = rng.random((64, N)).astype(dtype); C0 = rng.random((150, N)).astype(dtype)
    panels = sh.column_panels(N, world, 16)
    n0, wd = panels[rank]
    dB, dC = panel(B, n0, wd), panel(C0, n0, wd)
    op = xs.Fsspmdm(a, wd, beta=beta)
    op.execute_stream(dB.data_ptr(), dC.data_ptr())
    xs.synchronize(); xs.check()
    full = sh.gather_columns(dist, dC, panels, 150, tdt).cpu().numpy()
    op.destroy()
    if 0 == rank:
        want = C0.copy()
        if dtype == np.float64:
            orc.dfsspmdm_execute(a, B, want, beta, orc.dfsspmdm_branch(a, N, N, beta))
        else:
            orc.sfsspmdm_execute(a, B, want, beta)
        iv = np.uint64 if dtype == np.float64 else np.uint32
        assert np.array_equal(full.view(iv), want.view(iv)), "fsspmdm %%s beta=%%g: gathered C differs from the oracle" %% (dtype.__name__, beta)

# ---- spmdm: every rank slices its own replica of A and multiplies its panel (whole reference blocks of 48 columns) ----
for dt in ("f32", "bf16"):
    M, K, Nt = 512, 384, 48 * (7 * world + 1)
    A, Bf, Cf = w.spmdm_inputs(M, Nt, K, 0.05, dtype=dt, seed=3)
    panels = sh.column_panels(Nt, world, 48)
    n0, wd = panels[rank]
    p = xs.Spmdm(M, wd, K, 1)
    dA = torch.from_numpy(A).to(dev); dB, dC = panel(Bf, n0, wd), panel(Cf, n0, wd)
    beta = 0 if dt == "bf16" else 0.5
    p.create_slices(dA.data_ptr(), "N", dt == "bf16")
    p.compute(dB.data_ptr(), dC.data_ptr(), "N", "N", beta, dt == "bf16")
    xs.synchronize(); xs.check()
    geo = p.geometry
    full = sh.gather_columns(dist, dC, panels, M, torch.float32).cpu().numpy()
    p.destroy()
    if 0 == rank:
        gf = orc.geometry(M, Nt, K, 1, bn=48); gf.update(bm=geo["bm"], mb=geo["mb"])
        want = Cf.copy(); orc.compute(gf, orc.slices(gf, A), Bf, want, beta=float(beta))
        assert np.array_equal(full.view(np.uint32), want.view(np.uint32)), "spmdm %%s: gathered C differs from the oracle" %% dt
dist.barrier(); dist.destroy_process_group()
print("rank %%d ok on cuda:%%d" %% (rank, local))
'''


@pytest.mark.parametrize("world", [2, 4, 8])
def test_column_sharding_over_nccl(gpu, tmp_path, world):
    if gpu.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, gpu.device_count()))
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    for r in range(world):
        assert "rank %d ok" % r in out.stdout
