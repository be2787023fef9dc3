"""Multi-GPU partitioning of the sparse-A x dense-B path: by column panels of B and C, A replicated.

Column j of C depends only on column j of B and on A, so ranks never exchange data on the hot path
(SURVEY.md section 8e); a collective is used only to gather C for validation and to take the
max-over-ranks of the device time.  Host-side logic only (no CUDA): covered by the gloo tests.
"""


def column_panels(n_cols, world, multiple=16):
    """Splits [0, n_cols) into `world` contiguous panels whose widths are multiples of `multiple`
    (libxsmm_[sd]fsspmdm needs N % 16 == 0, reference src/libxsmm_fsspmdm.c:65) except possibly the last
    one, as evenly as possible.  Returns [(start, width)] of length `world` (width may be 0)."""
    if world < 1 or n_cols < 0 or multiple < 1:
        raise ValueError("column_panels(%r, %r, %r)" % (n_cols, world, multiple))
    units = n_cols // multiple
    rem = n_cols - units * multiple
    base, extra = divmod(units, world)
    out, start = [], 0
    for r in range(world):
        w = (base + (1 if r < extra else 0)) * multiple
        if r == world - 1:
            w += rem
        out.append((start, w))
        start += w
    assert start == n_cols
    return out


def panel_of(rank, n_cols, world, multiple=16):
    return column_panels(n_cols, world, multiple)[rank]


def gather_columns(dist, local_c, panels, n_rows, dtype):
    """all_gather of the per-rank C panels (validation only).  local_c: torch tensor n_rows x width."""
    import torch
    wmax = max(w for _, w in panels)
    pad = torch.zeros((n_rows, wmax), dtype=dtype, device=local_c.device)
    pad[:, :local_c.shape[1]] = local_c
    bufs = [torch.empty_like(pad) for _ in panels]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:, :w] for b, (_, w) in zip(bufs, panels)], dim=1)


def max_over_ranks(dist, value):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
