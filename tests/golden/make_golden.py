"""Generates the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref, built by oracle/build_ref.sh from /root/reference) on seeded inputs.

    python tests/golden/make_golden.py

The fixtures are what lets the CPU oracle (oracle/*.c) and the CUDA path be checked on machines where
/root/reference does not exist (the GPU box).  Inputs are stored next to the outputs so that nothing
depends on a PRNG stream being reproducible.  Files:
  geometry.json            libxsmm_spmdm_init geometry for a table of (M, N, K, max_threads)
  spmdm_<name>.npz         A, B, C0, geometry, slices (rowidx/colidx/values), C  for small spmdm cases
  fsspmdm_<name>.npz       a, B, C0, branch taken, chunk, C for beta 0 and 1 (synthetic operators)
  pyfr_<name>.npz          the same for real PyFR operators from samples/pyfr/mats/**/*.mtx
  fsspmdm_branch.json      which branch the reference's create() took for a sweep of operators
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402

w = importlib.import_module("libxsmm-1_b200.workloads")
REF_MATS = "/root/reference/samples/pyfr/mats"


def main():
    pyoracle.build_ref("avx2")
    ref = pyoracle.Ref()

    # ---- geometry table (reference src/libxsmm_spmdm.c:552-608) --------------------------------------
    table = []
    for (M, N, K) in [(2048, 2048, 2048), (4096, 4096, 4096), (1000, 1000, 1000), (300, 200, 260), (1, 16, 1),
                      (1024, 64, 4096), (1025, 48, 128), (4095, 100, 100), (8192, 512, 256), (33, 7, 129)]:
        for T in (1, 2, 8, 16, 56, 64, 96, 128, 192):
            g = ref.geometry(M, N, K, T)
            table.append(dict(M=M, N=N, K=K, T=T, **{k: g[k] for k in ("bm", "bn", "bk", "mb", "nb", "kb", "scratch")}))
    json.dump(table, open(os.path.join(HERE, "geometry.json"), "w"), indent=0)

    # ---- spmdm --------------------------------------------------------------------------------------
    cases = [
        ("f32_nnn_b0", 96, 100, 160, 0.15, "f32", "N", "N", "N", 0.0, 1),
        ("f32_nnn_b1", 96, 100, 160, 0.15, "f32", "N", "N", "N", 1.0, 1),
        ("f32_nnn_bhalf", 96, 103, 160, 0.15, "f32", "N", "N", "N", 0.5, 2),
        ("f32_tnt_b0", 70, 100, 140, 0.5, "f32", "T", "N", "T", 0.0, 1),
        ("f32_ntn_b0", 70, 100, 140, 0.5, "f32", "N", "T", "N", 0.0, 1),
        ("f32_ttt_b1", 70, 55, 140, 0.5, "f32", "T", "T", "T", 1.0, 1),
        ("bf16_nnn_b0", 96, 100, 256, 0.1, "bf16", "N", "N", "N", 0, 1),
        ("bf16_nnn_b1", 96, 100, 256, 0.1, "bf16", "N", "N", "N", 1, 1),
        ("bf16_tnt_b0", 64, 100, 128, 0.3, "bf16", "T", "N", "T", 0, 1),
        ("bf16_ntn_b0", 64, 100, 128, 0.3, "bf16", "N", "T", "N", 0, 1),
    ]
    for name, M, N, K, d, dt, ta, tb, tc, beta, T in cases:
        A, B, C0 = w.spmdm_inputs(M, N, K, d, dtype=dt, seed=len(name) + M, transa=ta, transb=tb, transc=tc)
        C = C0.copy()
        g, sl, _ = ref.spmdm(A, B, C, M, N, K, ta, tb, tc, beta, threads=T)
        np.savez_compressed(os.path.join(HERE, "spmdm_%s.npz" % name), A=A, B=B, C0=C0, C=C,
                            geom=np.array([g[k] for k in ("m", "n", "k", "bm", "bn", "bk", "mb", "nb", "kb")], np.int32),
                            rowidx=sl[0], colidx=sl[1], values=sl[2], trans=np.array([ta, tb, tc]),
                            beta=np.float64(beta), threads=np.int32(T))
    # special values (quirk Q3): NaN / -0 / Inf / denormal in the vector part and in the scalar remainder
    M, N, K = 40, 48, 133
    rng = np.random.default_rng(7)
    A = np.where(rng.random((M, K)) < 0.3, rng.random((M, K)), 0).astype(np.float32)
    A[0, 0] = np.nan; A[1, 5] = -0.0; A[2, 9] = np.inf; A[3, 11] = 1e-45; A[4, 127] = -np.inf
    A[5, 130] = np.nan; A[6, 131] = -0.0; A[7, 132] = np.inf
    B = rng.random((K, N)).astype(np.float32); C0 = rng.random((M, N)).astype(np.float32); C = C0.copy()
    g, sl, _ = ref.spmdm(A, B, C, M, N, K, beta=0.0)
    np.savez_compressed(os.path.join(HERE, "spmdm_f32_special.npz"), A=A, B=B, C0=C0, C=C,
                        geom=np.array([g[k] for k in ("m", "n", "k", "bm", "bn", "bk", "mb", "nb", "kb")], np.int32),
                        rowidx=sl[0], colidx=sl[1], values=sl[2], trans=np.array(["N", "N", "N"]), beta=np.float64(0), threads=np.int32(1))

    # ---- fsspmdm, synthetic -------------------------------------------------------------------------
    for name, dt, nu in [("d_sparse8", np.float64, 8), ("d_sparse31", np.float64, 31), ("d_dense32", np.float64, 32),
                         ("d_dense_cont", np.float64, None), ("s_dense8", np.float32, 8)]:
        a = w.fsspmdm_operator(48, 24, 0.3, nu, dt, seed=17)
        a[5, :] = 0     # an empty row: untouched by the sparse branch, zeroed by the dense one (quirk Q10)
        rng = np.random.default_rng(18)
        B = rng.random((24, 64)).astype(dt); C0 = rng.random((48, 64)).astype(dt)
        out = {}
        for beta in (0.0, 1.0):
            C = C0.copy()
            sparse, chunk, _ = ref.fsspmdm(a, B, C, beta, panel=64)
            out["C_beta%d" % int(beta)] = C
        np.savez_compressed(os.path.join(HERE, "fsspmdm_%s.npz" % name), a=a, B=B, C0=C0, sparse_branch=np.int32(sparse),
                            chunk=np.int32(chunk), **out)

    # ---- fsspmdm, real PyFR operators ------------------------------------------------------------------
    for rel in ["p3/hex/m6-sp.mtx", "p4/hex/m0-sp.mtx", "p4/tet/m6-sp.mtx", "p2/quad/m3-sp.mtx", "p3/tri/m0-sp.mtx"]:
        path = os.path.join(REF_MATS, rel)
        if not os.path.exists(path):
            continue
        a = w.read_mtx(path)
        rng = np.random.default_rng(19)
        B = rng.random((a.shape[1], 32)); C0 = rng.random((a.shape[0], 32))
        out = {}
        for beta in (0.0, 1.0):
            C = C0.copy()
            sparse, chunk, _ = ref.fsspmdm(a, B, C, beta, panel=32)
            out["C_beta%d" % int(beta)] = C
        np.savez_compressed(os.path.join(HERE, "pyfr_%s.npz" % rel.replace("/", "_").replace("-sp.mtx", "")), a=a, B=B, C0=C0,
                            sparse_branch=np.int32(sparse), chunk=np.int32(chunk), **out)

    # ---- branch rule sweep (quirk Q9): unique-value count and the 128 KiB code limit -----------------------
    sweep = []
    for nu in (1, 2, 30, 31, 32, 33, 64, None):
        a = w.fsspmdm_operator(40, 32, 0.4, nu, np.float64, seed=23)
        B = np.zeros((32, 16)); C = np.zeros((40, 16))
        sparse, chunk, _ = ref.fsspmdm(a, B, C, 0.0, panel=16)
        sweep.append(dict(kind="unique", M=40, K=32, density=0.4, n_unique=nu, seed=23, ld=16, beta=0.0, sparse=bool(sparse), chunk=chunk))
    # code size: a 2-value operator whose emitted kernel crosses 131072 bytes as it grows
    for M in (400, 600, 700, 800, 900, 1000, 1200):
        for ld in (16, 4096):
            for beta in (0.0, 1.0):
                a = w.fsspmdm_operator(M, 16, 0.9, 2, np.float64, seed=29)
                B = np.zeros((16, ld)); C = np.zeros((M, ld))
                sparse, chunk, _ = ref.fsspmdm(a, B, C, beta, N=16, ld=ld, panel=16)
                sweep.append(dict(kind="codesize", M=M, K=16, density=0.9, n_unique=2, seed=29, ld=ld, beta=beta, sparse=bool(sparse), chunk=chunk))
    json.dump(sweep, open(os.path.join(HERE, "fsspmdm_branch.json"), "w"), indent=0)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
