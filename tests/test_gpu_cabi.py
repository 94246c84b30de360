"""The drop-in boundary without Python in the loop: tests/c/abi_smoke.c (plain C, only include/oetqf_b200.h) is
compiled with gcc, linked against liboetqf_b200.so and run; what it computed is compared with the same problem
evaluated by the CPU oracle.  This is the closest executable stand-in for the reference's `ccall` binding
(/root/reference/src/pref.jl:1-21, src/BEM/equation.jl:141-154) in an environment without Julia."""
import os
import struct
import subprocess

import numpy as np
import pytest

import workloads as W
from helpers import scaled_err
from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
LIBDIR = os.path.join(ROOT, "oetqf.jl_b200")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           SRC, "-o", exe, "-L", LIBDIR, "-loetqf_b200", "-lm", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def _read(path):
    out = {}
    with open(path, "rb") as fh:
        while True:
            tag = fh.read(32)
            if len(tag) < 32:
                break
            (n,) = struct.unpack("<Q", fh.read(8))
            out[tag.rstrip(b"\0").decode()] = np.frombuffer(fh.read(8 * n), dtype=np.float64).copy()
    return out


def test_c_program_compiles_against_the_header_alone(tmp_path):
    """CPU-side half (also run by -m "not gpu" through test_abi): the header is valid C11, the _Static_asserts on
    the struct layouts hold, every symbol the program uses resolves at link time"""
    _build(tmp_path)


@pytest.mark.gpu
def test_c_program_end_to_end(gpu, tmp_path):
    exe = _build(tmp_path)
    out = str(tmp_path / "abi.bin")
    res = subprocess.run([exe, out], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, (res.stdout, res.stderr)
    assert "abi_smoke ok" in res.stdout
    d = _read(out)
    fs, bs = W.C2_FAULT, W.C2_BOX
    mfo = ref.fault_mesh(fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    mao = ref.hex8_box(*bs.args())
    nf, ne = 32, 36
    o11 = ref.gf_fault_fault(mfo, W.LAM, W.MU, buffer_ratio=1.0)
    o12 = ref.gf_fault_mantle(mfo, mao, W.LAM, W.MU, buffer_ratio=1.0)
    o21 = ref.gf_mantle_fault(mao, mfo, W.LAM, W.MU)
    o22 = ref.gf_mantle_mantle(mao, W.LAM, W.MU)
    # the C program generated the same meshes on its own: Okada entries are bit-identical, hex8 to 1e-10 of the scale
    assert np.array_equal(d["g11"], o11.reshape(-1, order="F"))
    assert np.array_equal(d["g12"], o12.reshape(-1, order="F"))
    assert scaled_err(d["g21"].reshape(o21.shape, order="F"), o21, axis=0) < 1e-10
    assert scaled_err(d["g22"].reshape(o22.shape, order="F"), o22, axis=0) < 1e-10
    xv = np.sin(0.37 * np.arange(6 * ne)) * 1e-14
    want = ref.gemv(o21, xv)
    assert np.max(np.abs(d["gemv21"] - want)) <= 1e-10 * np.max(np.abs(want))
    want22 = ref.gemv(o22, xv)                       # the class-form operand (oq_matrix_mantle_mantle_classes) through oq_gemv
    assert np.max(np.abs(d["gemv22_classes"] - want22)) <= 1e-10 * np.max(np.abs(want22))
    # the (du, u, p, t) call
    shp = (mfo.nx, mfo.nxi)
    pf = ref.FaultProp(*(d[k].reshape(shp, order="F") for k in ("a", "b", "L", "sigma")), W.ETA, W.VPL, 0.6, 1e-6)
    pa = ref.MantleProp(d["gamma"], np.full(ne, 2.5), W.DEPS0)
    v, th = d["u_v"].reshape(shp, order="F"), d["u_th"].reshape(shp, order="F")
    sig = d["u_sig"].reshape((ne, 6), order="F")
    dv, dth, deps, dsig, ddl = ref.rhs_viscoelastic(pf, pa, o11, o12, o21, o22, v, th, sig, form="toeplitz")
    for got, w in ((d["dv"], dv), (d["dth"], dth), (d["deps"], deps), (d["dsig"], dsig), (d["ddl"], ddl)):
        w = np.asarray(w).reshape(-1, order="F")
        den = np.maximum(np.abs(w), 1e-6 * np.max(np.abs(w)) + 1e-300)
        assert np.max(np.abs(got - w) / den) < 1e-9
    t, nacc, nrej, nrhs, retcode, nsnap = d["stats"]
    assert retcode == 0 and nsnap == nacc + 1 and abs(t - 1e-3 * W.YEAR) < 1e-6
    assert np.all(np.isfinite(d["fin_v"])) and np.all(d["fin_th"] > 0)
