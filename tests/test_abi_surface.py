"""CPU-only checks of the drop-in boundary: libellcuda.so loads without a GPU, exports every entry point that
include/ellcuda.h declares, its records have the layout the ctypes mirror assumes, and compute calls fail loudly
(no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from ellalgo_rs_b200 import _abi as A
from ellalgo_rs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ellcuda.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ellcuda_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    _lib.build()
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in ellcuda.h but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == names, "ellalgo_rs_b200/_lib.py EXPORTS out of sync with the header"


def test_struct_layout_matches_the_header(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ellcuda.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(ell_batch_desc), sizeof(ell_batch_result), sizeof(ell_create_desc), offsetof(ell_batch_desc, table),"
                   "offsetof(ell_batch_desc, stream), offsetof(ell_create_desc, stream));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(A.BatchDesc), C.sizeof(A.BatchResult), C.sizeof(A.CreateDesc), A.BatchDesc.table.offset,
            A.BatchDesc.stream.offset, A.CreateDesc.stream.offset]
    assert got == want


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only box")
    from ellalgo_rs_b200 import batch, workloads as W
    from ellalgo_rs_b200.ell import Ell
    assert _lib.lib().ellcuda_device_count() == 0
    with pytest.raises(_lib.EllCudaError) as e:
        batch.solve_batch(W.example1())
    assert e.value.code == A.E_CUDA
    with pytest.raises(_lib.EllCudaError):
        Ell.new_with_scalar(1.0, np.zeros(4))
    x = np.zeros(4)
    assert _lib.lib().ellcuda_math_probe(0, x.ctypes.data, x.ctypes.data, 4, 0) == A.E_CUDA


def test_host_side_table_needs_no_gpu(ref):
    from ellalgo_rs_b200 import batch
    assert np.array_equal(batch.lowpass_table(32), ref.lowpass_table(32))   # LowpassOracle::new's spectrum, glibc cos
    assert _lib.lib().ellcuda_lowpass_band_edge(32, 0.12) == ref.lib().ellref_lowpass_band_edge(32, 0.12) == 58


def test_product_never_imports_the_checker():
    pkg = os.path.join(ROOT, "ellalgo_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ellref" not in text and "oracle/" not in text, f"{f} refers to the checker"
