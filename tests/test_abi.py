"""The C-ABI boundary: libxsmm_b200.so loads without a GPU and exports every symbol include/*.h
declares; the public structs have the reference's layout (reference include/libxsmm_spmdm.h:42-71,
.abi.txt:47-49,352-354,376-383)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"LIBXSMM_API\s+[^;(]*?\b(libxsmm_\w+)\s*\(", text):
            names.add(m.group(1))
    return names


def test_exports_every_declared_symbol(xs):
    decl = declared_symbols()
    assert set(xs.REFERENCE_SYMBOLS) <= decl
    assert len(decl) >= 14 + 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", xs.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = decl - exported
    assert not missing, missing
    lib = xs.load()
    for name in decl:
        assert getattr(lib, name) is not None
    # nothing but the declared interface leaks out of the library
    stray = {s for s in exported if not s.startswith("libxsmm_")}
    assert not stray, stray
    assert set(xs.REFERENCE_SYMBOLS + xs.ADDED_SYMBOLS) == decl


def test_struct_layout(xs):
    h = xs.libxsmm_spmdm_handle
    assert ctypes.sizeof(h) == 64
    assert h.base_ptr_scratch_A.offset == 40 and h.base_ptr_scratch_B_scratch_C.offset == 48
    assert h.memory_for_scratch_per_thread.offset == 56 and h.datatype.offset == 36
    assert ctypes.sizeof(xs.libxsmm_CSR_sparseslice) == 24


def test_headers_compile_as_c89_and_cpp():
    src = '#include "libxsmm_b200.h"\nint main(void) { libxsmm_spmdm_handle h; (void)h; return (int)sizeof(libxsmm_CSR_sparseslice) - 24; }\n'
    for cc, flags in (("gcc", ["-x", "c", "-std=c89", "-Wall", "-Werror"]), ("g++", ["-x", "c++", "-Wall", "-Werror"])):
        subprocess.run([cc] + flags + ["-I", os.path.join(ROOT, "include"), "-fsyntax-only", "-"], input=src, text=True, check=True)


def test_no_compute_without_gpu_is_loud(xs):
    """no CPU fallback: with no device the constructors raise instead of computing elsewhere."""
    if xs.device_count() > 0:
        return
    import numpy as np
    import pytest
    with pytest.raises(RuntimeError):
        xs.Spmdm(64, 64, 64)
    with pytest.raises(RuntimeError):
        xs.Fsspmdm(np.eye(16), 16)
    h, s = xs.libxsmm_spmdm_init(64, 64, 64, 1)      # the raw entry mirrors the reference: NULL arenas, error recorded
    assert not h.base_ptr_scratch_A
    assert xs.last_error()[0] != 0
    xs.clear_error()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "libxsmm-1_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text, os.path.join(dirpath, f)
