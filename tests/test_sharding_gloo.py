"""N > 1 path on CPU: world_size-2 gloo.  The data path has no collective; what is covered here is the
column partition, the per-rank panel views (row pitch = full N for 'N' layouts), the gather used for
validation and the max-over-ranks timing reduction.  Each rank's panel is produced by the CPU oracle
(the checker), standing in for the device, and the gathered C must equal the oracle's full C bit for bit
because columns are independent."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import pyoracle
xs_sh = importlib.import_module("libxsmm-1_b200.sharding")
w = importlib.import_module("libxsmm-1_b200.workloads")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
orc = pyoracle.Oracle()
# ---- fsspmdm: N columns split in multiples of 16 ---------------------------------------------------------
a = w.fsspmdm_operator(40, 24, 0.3, 6, np.float64, seed=1)
N = 16 * 13
rng = np.random.default_rng(2)
B = rng.random((24, N)); C0 = rng.random((40, N))
panels = xs_sh.column_panels(N, world, 16)
n0, wd = panels[rank]
Bl = np.ascontiguousarray(B[:, n0:n0 + wd]); Cl = np.ascontiguousarray(C0[:, n0:n0 + wd])
orc.dfsspmdm_execute(a, Bl, Cl, 1.0, orc.dfsspmdm_branch(a, wd, wd, 1.0))
full = xs_sh.gather_columns(dist, torch.from_numpy(Cl), panels, 40, torch.float64).numpy()
want = C0.copy(); orc.dfsspmdm_execute(a, B, want, 1.0, orc.dfsspmdm_branch(a, N, N, 1.0))
assert np.array_equal(full.view(np.uint64), want.view(np.uint64)), "fsspmdm gather mismatch"
# ---- spmdm: each rank owns a panel; geometry per rank uses N_r -----------------------------------------------
M, K, Nt = 96, 160, 96 * world          # panels of 96 columns = whole reference blocks (bn = 48): same rounding as the full problem
A, Bf, Cf = w.spmdm_inputs(M, Nt, K, 0.2, seed=3)
panels = xs_sh.column_panels(Nt, world, 48)
n0, wd = panels[rank]
g = orc.geometry(M, wd, K, 1, bn=48)
sl = orc.slices(g, A)
Cl = np.ascontiguousarray(Cf[:, n0:n0 + wd]); orc.compute(g, sl, np.ascontiguousarray(Bf[:, n0:n0 + wd]), Cl, beta=0.5)
full = xs_sh.gather_columns(dist, torch.from_numpy(Cl), panels, M, torch.float32).numpy()
gf = orc.geometry(M, Nt, K, 1, bn=48); gf.update(bm=g.bm, mb=g.mb)
want = Cf.copy(); orc.compute(gf, orc.slices(gf, A), Bf, want, beta=0.5)
assert np.array_equal(full.view(np.uint32), want.view(np.uint32)), "spmdm gather mismatch"
# ---- timing reduction -----------------------------------------------------------------------------------------
assert xs_sh.max_over_ranks(dist, 1.0 + rank) == float(world)
dist.barrier(); dist.destroy_process_group()
print("rank %%d ok" %% rank)
'''


def test_column_panels():
    from importlib import import_module
    sh = import_module("libxsmm-1_b200.sharding")
    for n in (0, 16, 48, 1000, 1 << 20, (1 << 24) + 7):
        for world in (1, 2, 3, 4, 8):
            p = sh.column_panels(n, world, 16)
            assert len(p) == world and p[0][0] == 0
            assert sum(w for _, w in p) == n
            for (s0, w0), (s1, _) in zip(p, p[1:]):
                assert s0 + w0 == s1 and w0 % 16 == 0
            widths = [w for _, w in p[:-1]]
            assert not widths or max(widths) - min(widths) <= 16
    assert sh.column_panels(1 << 20, 8, 16) == [(i << 17, 1 << 17) for i in range(8)]
    with pytest.raises(ValueError):
        sh.column_panels(16, 0)


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
