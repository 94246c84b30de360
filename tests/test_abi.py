"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol the header declares,
and fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, "include", "oetqf_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(oq_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(oq):
    assert _header_symbols() == sorted(oq._lib.EXPORTS)


def test_library_exports_every_declared_symbol(oq):
    lib = oq._lib.load()
    for name in _header_symbols():
        assert hasattr(lib, name), f"{name} declared in include/oetqf_b200.h but not exported"
    assert lib.oq_abi_version() == 2


def test_no_cpu_fallback(oq):
    n = ctypes.c_int(0)
    oq._lib.load().oq_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    mf = oq.gen_mesh("RectOkada", 100.0, 100.0, 10.0, 10.0, 41.0)
    with pytest.raises(oq.OqError, match="no CUDA device|no CPU fallback"):
        oq.stress_greens_function(mf, 3e10, 3e10, fourier=False)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package, the examples, the helper scripts, the
    workload definitions or the import shim may import, include, link or load it (only tests/, smoke() and the CPU
    legs of bench.py do)."""
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s+[\"<][^\">]*oracle)|liboetqf_oracle|oracle\.ref|oracle/ref", re.M)
    roots = [os.path.join(ROOT, d) for d in ("oetqf.jl_b200", "examples", "scripts", "include")]
    singles = [os.path.join(ROOT, f) for f in ("workloads.py", "oetqf_b200.py")]
    for top in roots:
        for base, _, files in os.walk(top):
            if os.path.basename(base) == "derive":
                continue      # the generator WRITES the checker's copy of the closed form; it never reads the oracle
            singles += [os.path.join(base, f) for f in files
                        if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", ".sh", "Makefile"))]
    for path in singles:
        with open(path, errors="replace") as fh:
            assert not bad.search(fh.read()), f"{path} reaches into oracle/"


def test_argument_validation_precedes_device_use(oq):
    """every export returns non-zero with a message instead of crashing on NULL / malformed arguments
    (checked without a GPU: validation happens before the device is touched)"""
    import ctypes as C
    lib = oq._lib.load()
    d = C.c_double(0)
    assert lib.oq_gf_fault_fault(None, C.c_double(1), C.c_double(1), 0, 0, 2, C.c_double(0), None, None) != 0
    assert b"NULL" in lib.oq_last_error()
    assert lib.oq_gemv(None, None, None, 0) != 0
    assert lib.oq_matrix_to_host(None, None) != 0
    assert lib.oq_problem_layout(None, None, None) != 0
    assert lib.oq_solve(None, C.c_double(0), None, 1, C.cast(None, oq._lib.SNAPSHOT_FN), None, None) != 0
    assert lib.oq_dc3d_gradient(-1, None, None, None, d, d, d, d, d, d, d, 0, None) != 0
    assert lib.oq_dc3d_gradient(0, None, None, None, d, d, d, d, d, d, d, 7, None) != 0
    assert b"fault type" in lib.oq_last_error()
    assert lib.oq_matrix_destroy(None) == 0 and lib.oq_problem_destroy(None) == 0     # destroying NULL is a no-op
