"""Seeded synthetic inputs for the SPMDM / FSSPMDM configurations.

The PRNG is harness-owned (numpy PCG64) -- deliberately NOT libxsmm_rng_f64, which is
libc dependent (reference src/libxsmm_rng.c:250-258).  Values are uniform in [0,1) like the
reference samples (samples/spmdm/spmdm.c:212-243); the mask is an independent draw.
"""
import numpy as np


def to_bf16_bits(x):
    """fp32 -> bf16 by truncation (top 16 bits), as samples/spmdm/spmdm.c:217-218 does."""
    x = np.ascontiguousarray(x, np.float32)
    return (x.view(np.uint32) >> np.uint32(16)).astype(np.uint16)


def from_bf16_bits(h):
    return (np.ascontiguousarray(h, np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def spmdm_inputs(M, N, K, density, dtype="f32", seed=1, transa="N", transb="N", transc="N", as_shipped=False):
    """Returns (A, B, C0).  A is stored M x K ('N') or K x M ('T'); B K x N or N x K; C M x N or N x M.
    dtype 'f32' -> float32 arrays; 'bf16' -> uint16 bit patterns for A and B, float32 C."""
    rng = np.random.default_rng(seed)
    if as_shipped:  # one draw for mask and value, threshold 0.85 (samples/spmdm/spmdm.c:212-224)
        r = rng.random((M, K), np.float32)
        A = np.where(r > np.float32(1.0 - density), r, np.float32(0))
    else:
        u = rng.random((M, K), np.float32)
        v = rng.random((M, K), np.float32)
        v[v == 0] = np.float32(0.5)
        A = np.where(u < np.float32(density), v, np.float32(0))
    B = rng.random((K, N), np.float32)
    C = rng.random((M, N), np.float32)
    if dtype == "bf16":
        A = from_bf16_bits(to_bf16_bits(A))
        # truncation may produce zeros from tiny values; keep them (they are simply dropped)
        B = from_bf16_bits(to_bf16_bits(B))
    if transa in "Tt":
        A = np.ascontiguousarray(A.T)
    if transb in "Tt":
        B = np.ascontiguousarray(B.T)
    if transc in "Tt":
        C = np.ascontiguousarray(C.T)
    if dtype == "bf16":
        return to_bf16_bits(A), to_bf16_bits(B), np.ascontiguousarray(C)
    return np.ascontiguousarray(A), np.ascontiguousarray(B), np.ascontiguousarray(C)


def fsspmdm_operator(M=150, K=64, density=0.30, n_unique=8, dtype=np.float64, seed=1, no_empty_rows=True):
    """PyFR-style fixed operator.  n_unique = size of the value pool (<= 31 makes the reference
    take its sparse_reg branch for double); None = continuous values (reference dense branch)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((M, K)) < density
    if no_empty_rows:
        for i in np.nonzero(mask.sum(1) == 0)[0]:
            mask[i, rng.integers(K)] = True
    if n_unique is None:
        vals = rng.random((M, K)) + 0.25
    else:
        pool = (np.arange(1, n_unique + 1) / float(n_unique + 1)) * np.where(np.arange(n_unique) % 2, -1.0, 1.0)
        vals = pool[rng.integers(0, n_unique, (M, K))]
    return np.ascontiguousarray(np.where(mask, vals, 0.0).astype(dtype))


def read_mtx(path, dtype=np.float64):
    """Coordinate MatrixMarket reader for the PyFR operator files (1-based, row-sorted;
    the reference harness's own reader is samples/pyfr/pyfr_driver_asp_reg.c:47-158).
    Returns the dense M x K operator."""
    with open(path, "r") as f:
        header = None
        rows = []
        for line in f:
            if line.startswith("%"):
                continue
            parts = line.split()
            if header is None:
                header = (int(parts[0]), int(parts[1]), int(parts[2]))
                continue
            rows.append((int(parts[0]) - 1, int(parts[1]) - 1, float(parts[2])))
    M, K, nnz = header
    A = np.zeros((M, K), dtype)
    for i, j, v in rows:
        A[i, j] = v
    return A
