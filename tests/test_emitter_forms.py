"""The PTX emitter of the baked fsspmdm kernels, checked without a GPU: for fma-heavy fp64 operators create() bakes up to three
forms (plain; rows interleaved explicitly 8 or 2 at a time with the operator's values in constant memory) and keeps the
fastest.  Here every form of a few real PyFR operators (tests/golden/operators_pyfr.npz) is emitted through
libxsmm_b200_fsspmdm_kernel_source, must hold exactly one fma per nonzero and per row one store, and must assemble for sm_100a
with the toolkit's ptxas (the driver's assembler does the same at create)."""
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PTXAS = shutil.which("ptxas") or "/usr/local/cuda/bin/ptxas"
FORMS = {"plain": dict(IL="0", CONST="0", MINCTAS="0"), "il8": dict(IL="8", CONST="1", MINCTAS="3"), "il2": dict(IL="2", CONST="1", MINCTAS="3")}


def operator(name):
    d = np.load(os.path.join(ROOT, "tests", "golden", "operators_pyfr.npz"))
    i = [str(n) for n in d["names"]].index(name)
    M, K = (int(x) for x in d["shapes"][i])
    lo, hi = int(d["offsets"][i]), int(d["offsets"][i + 1])
    a = np.zeros((M, K))
    a[d["rows"][lo:hi].astype(np.int64), d["cols"][lo:hi].astype(np.int64)] = d["vals"][lo:hi]
    return a


@pytest.mark.parametrize("name", ["p4/tet/m6", "p4/tet/m3", "p6/tri/m132"])
@pytest.mark.parametrize("form", sorted(FORMS))
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_form_is_complete_and_assembles(xs, monkeypatch, name, form, beta):
    if not os.path.exists(PTXAS):
        pytest.skip("ptxas not installed")
    for k, v in FORMS[form].items():
        monkeypatch.setenv("LIBXSMM_B200_FSSPMDM_" + k, v)
    a = operator(name)
    ptx = xs.fsspmdm_kernel_source(a, N=1 << 20, beta=beta)
    assert ptx and ".target sm_100a" in ptx
    nnz, rows = int(np.count_nonzero(a)), int(np.count_nonzero(np.count_nonzero(a, axis=1)))
    assert len(re.findall(r"fma\.rn\.f64", ptx)) == nnz              # one fma per nonzero in every form
    assert len(re.findall(r"st\.global\.cs\.f64", ptx)) == rows        # every non-empty row stored once
    explicit = form != "plain"
    assert (".const .align 8 .b64 fsv[" in ptx) == explicit
    assert ("bar.warp.sync" in ptx) == explicit
    assert (".minnctapersm 3" in ptx) == explicit
    if explicit:      # every fma takes its value from the table, whose entries are the operator's distinct values
        assert len(re.findall(r"ld\.const\.f64", ptx)) == nnz
        table = re.search(r"fsv\[(\d+)\]", ptx)
        assert int(table.group(1)) == len(np.unique(a[a != 0]))
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "k.ptx")
        open(src, "w").write(ptx)
        r = subprocess.run([PTXAS, "-arch=sm_100a", src, "-o", os.path.join(tmp, "k.cubin")], capture_output=True, text=True)
        assert 0 == r.returncode, r.stderr[:2000]


def test_each_rows_own_order_is_kept(xs, monkeypatch):
    """the explicit forms interleave the chains of several rows but never reorder a row's own fmas: per accumulator the sequence
    of B registers is the row's ascending column order"""
    a = operator("p4/tet/m3")
    for k, v in FORMS["il8"].items():
        monkeypatch.setenv("LIBXSMM_B200_FSSPMDM_" + k, v)
    ptx = xs.fsspmdm_kernel_source(a, N=1 << 20, beta=0.0)
    chains = {}
    order = []
    for line in ptx.splitlines():
        m = re.match(r"\s*mov\.f64 %ai(\d), 0d0000000000000000;", line)
        if m:
            chains[int(m.group(1))] = []
            continue
        m = re.match(r"\s*fma\.rn\.f64 %ai(\d), %kv, %bx(\d+), %ai\1;", line)
        if m:
            chains[int(m.group(1))].append(int(m.group(2)))
            continue
        m = re.match(r"\s*st\.global\.cs\.f64 \[%rd27\], %ai(\d);", line)
        if m:
            order.append(chains.pop(int(m.group(1))))
    rows = [list(np.nonzero(r)[0]) for r in a if np.count_nonzero(r)]
    assert order == rows
