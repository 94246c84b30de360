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
    assert lib.oq_abi_version() == 1


def test_no_cpu_fallback(oq):
    n = ctypes.c_int(0)
    oq._lib.load().oq_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    mf = oq.gen_mesh("RectOkada", 100.0, 100.0, 10.0, 10.0, 41.0)
    with pytest.raises(oq.OqError, match="no CUDA device|no CPU fallback"):
        oq.stress_greens_function(mf, 3e10, 3e10, fourier=False)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, include, link or load it."""
    pkg = os.path.join(ROOT, "oetqf.jl_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s+[\"<][^\">]*oracle)|liboetqf_oracle|oracle\.ref|oracle/ref", re.M)
    for base, _, files in os.walk(pkg):
        if os.path.basename(base) == "derive":
            continue          # the generator WRITES the checker's copy of the closed form; it never reads the oracle
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
                with open(os.path.join(base, f), errors="replace") as fh:
                    assert not bad.search(fh.read()), f"{f} reaches into oracle/"
